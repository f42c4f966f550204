// BatchNorm1d (+ReLU) over the live rows of a [capacity, C] feature buffer whose row count lives
// on the device (graph path).  These are the modules the reference's SparseSequential applies to
// `.features` between the sparse convolutions (src/models/SPConvBlocks.py:505-508).  Arithmetic
// follows torch.nn.BatchNorm1d in training mode: biased batch variance for the normalisation,
// unbiased variance for the running estimate, running = (1 - momentum) * running + momentum * new.
//
// Pure streaming work (HBM-bound): forward reads x twice (statistics, normalise) and writes y once;
// statistics are per-(row-chunk, channel) Welford partials (count, mean, M2) merged in double in a
// fixed order, so the result does not depend on scheduling.
#include "common.cuh"

namespace wfsp {
namespace {

constexpr int kRows = 128;  // rows per partial
constexpr int kCh = 32;     // channels per block (one 128-byte row segment)

__device__ __forceinline__ int64_t live_rows(int64_t n, const int32_t* n_dev) { return n_dev ? int64_t(*n_dev) : n; }

// grid (ceil(cap/kRows), ceil(c/kCh)), block (32, 8)
__global__ void __launch_bounds__(256) bn_partial_stats(const float* __restrict__ x, int64_t n_cap,
                                                        const int32_t* __restrict__ n_dev, int c,
                                                        float* __restrict__ part /* [nblk][2][c] mean, M2 */) {
  __shared__ float red[8][kCh];
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t r0 = int64_t(blockIdx.x) * kRows;
  if (r0 >= n) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.y * kCh + tx;
  const int rows = int(n - r0 < kRows ? n - r0 : kRows);
  float s = 0.f;
  if (ch < c)
    for (int r = ty; r < rows; r += 8) s += x[(r0 + r) * c + ch];
  red[ty][tx] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i][tx];
  const float mean = tot / float(rows);
  __syncthreads();
  float m2 = 0.f;
  if (ch < c)
    for (int r = ty; r < rows; r += 8) {
      const float d = x[(r0 + r) * c + ch] - mean;
      m2 += d * d;
    }
  red[ty][tx] = m2;
  __syncthreads();
  if (ty == 0 && ch < c) {
    float t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t2 += red[i][tx];
    part[(int64_t(blockIdx.x) * 2 + 0) * c + ch] = mean;
    part[(int64_t(blockIdx.x) * 2 + 1) * c + ch] = t2;
  }
}

// One block per 32 channels; 32 row lanes each merge every 32nd partial (Chan's parallel-variance merge, in
// double), then lane 0 merges the 32 lane results in lane order: a fixed order, so the statistics do
// not depend on scheduling, and nblk/32 dependent steps instead of nblk.
constexpr int kFinLanes = 32;

__global__ void __launch_bounds__(kCh * kFinLanes) bn_finalize_stats(const float* __restrict__ part, int64_t n_cap,
                                                         const int32_t* __restrict__ n_dev, int c, float eps,
                                                         float momentum, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var,
                                                         float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ double s_cnt[kFinLanes][kCh], s_mean[kFinLanes][kCh], s_m2[kFinLanes][kCh];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t nblk = n > 0 ? (n + kRows - 1) / kRows : 0;
  double cnt = 0.0, mean = 0.0, m2 = 0.0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t b = ty; b < nblk; b += kFinLanes) {
      const double nb = double(b + 1 < nblk ? kRows : n - b * kRows);
      const double mb = part[(b * 2 + 0) * c + ch], m2b = part[(b * 2 + 1) * c + ch];
      const double delta = mb - mean, tot = cnt + nb;
      mean += delta * nb / tot;
      m2 += m2b + delta * delta * cnt * nb / tot;
      cnt = tot;
    }
  }
  s_cnt[ty][tx] = cnt; s_mean[ty][tx] = mean; s_m2[ty][tx] = m2;
  __syncthreads();
  if (ty != 0 || ch >= c) return;
  if (n <= 0) { save_mean[ch] = 0.f; save_invstd[ch] = 0.f; return; }
  for (int l = 1; l < kFinLanes; ++l) {
    const double nb = s_cnt[l][tx];
    if (nb <= 0.0) continue;
    const double delta = s_mean[l][tx] - mean, tot = cnt + nb;
    mean += delta * nb / tot;
    m2 += s_m2[l][tx] + delta * delta * cnt * nb / tot;
    cnt = tot;
  }
  const double var = m2 / cnt;
  save_mean[ch] = float(mean);
  save_invstd[ch] = float(1.0 / sqrt(var + double(eps)));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * float(mean);
  if (running_var && cnt > 1.0) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * float(m2 / (cnt - 1.0));
}

__global__ void __launch_bounds__(128) bn_eval_stats(int c, float eps, const float* __restrict__ running_mean,
                                                     const float* __restrict__ running_var,
                                                     float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  save_mean[ch] = running_mean[ch];
  save_invstd[ch] = rsqrtf(running_var[ch] + eps);
}

__device__ __forceinline__ uint32_t bn_pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Streaming normalise(+ReLU) pass.  A block owns kApplyRows consecutive rows (one contiguous span of x);
// a thread handles two adjacent channels of a row at a time, so a warp reads 256 contiguous bytes and
// writes 256 (fp32) / 128 (bf16) contiguous bytes per instruction.  Outputs (each optional): y fp32
// [rows, c] and y16 bf16 [rows, c_pad] (c_pad = c rounded up to 8, padding written as zero) -- the operand
// format of the tcgen05 convolution kernels, so the next layer needs no cast pass.
// mean == nullptr switches the normalisation off (plain ReLU / cast).
constexpr int kApplyRows = 32;

__global__ void __launch_bounds__(256) bn_apply(const float* __restrict__ x, int64_t n_cap,
                                                const int32_t* __restrict__ n_dev, int c,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                int relu, float* __restrict__ y, __nv_bfloat16* __restrict__ y16) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;  // channel pairs per (padded) row
  const float inv_ppr = 1.f / float(ppr);
  for (int64_t r0 = int64_t(blockIdx.x) * kApplyRows; r0 < n; r0 += int64_t(gridDim.x) * kApplyRows) {
    const int rows = int(n - r0 < kApplyRows ? n - r0 : kApplyRows);
    const float* xb = x + r0 * c;
    float* yb = y ? y + r0 * c : nullptr;
    uint32_t* y16b = y16 ? reinterpret_cast<uint32_t*>(y16 + r0 * c_pad) : nullptr;
    const int total = rows * ppr;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      int row = int((float(t) + 0.5f) * inv_ppr);
      int pr = t - row * ppr;
      if (pr < 0) { --row; pr += ppr; } else if (pr >= ppr) { ++row; pr -= ppr; }
      const int ch = pr << 1;
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (ch + e < c) {
          float tv = xb[row * c + ch + e];
          if (mean) tv = (tv - mean[ch + e]) * invstd[ch + e] * (gamma ? gamma[ch + e] : 1.f) + (beta ? beta[ch + e] : 0.f);
          if (relu && tv < 0.f) tv = 0.f;
          if (yb) yb[row * c + ch + e] = tv;
          v[e] = tv;
        }
      }
      if (y16b) y16b[row * ppr + pr] = bn_pack_bf16x2(v[0], v[1]);
    }
  }
}

// backward partials: sum(dy') and sum(dy' * xhat) per (row chunk, channel); dy' = dy masked by relu
__global__ void __launch_bounds__(256) bn_bwd_partial(const float* __restrict__ x, const float* __restrict__ dy,
                                                      int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                      int relu, float* __restrict__ part /* [nblk][2][c] */) {
  __shared__ float red0[8][kCh], red1[8][kCh];
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t r0 = int64_t(blockIdx.x) * kRows;
  if (r0 >= n) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.y * kCh + tx;
  const int rows = int(n - r0 < kRows ? n - r0 : kRows);
  float s0 = 0.f, s1 = 0.f;
  if (ch < c) {
    const float m = mean[ch], is = invstd[ch], g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
    for (int r = ty; r < rows; r += 8) {
      const float xh = (x[(r0 + r) * c + ch] - m) * is;
      float d = dy[(r0 + r) * c + ch];
      if (relu && xh * g + b <= 0.f) d = 0.f;
      s0 += d;
      s1 += d * xh;
    }
  }
  red0[ty][tx] = s0;
  red1[ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && ch < c) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += red0[i][tx]; t1 += red1[i][tx]; }
    part[(int64_t(blockIdx.x) * 2 + 0) * c + ch] = t0;
    part[(int64_t(blockIdx.x) * 2 + 1) * c + ch] = t1;
  }
}

__global__ void __launch_bounds__(kCh * kFinLanes) bn_bwd_finalize(const float* __restrict__ part, int64_t n_cap,
                                                       const int32_t* __restrict__ n_dev, int c,
                                                       float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ double r0[kFinLanes][kCh], r1[kFinLanes][kCh];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t nblk = n > 0 ? (n + kRows - 1) / kRows : 0;
  double s0 = 0.0, s1 = 0.0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t b = ty; b < nblk; b += kFinLanes) {
      s0 += part[(b * 2 + 0) * c + ch];
      s1 += part[(b * 2 + 1) * c + ch];
    }
  }
  r0[ty][tx] = s0; r1[ty][tx] = s1;
  __syncthreads();
  if (ty != 0 || ch >= c) return;
  for (int l = 1; l < kFinLanes; ++l) { s0 += r0[l][tx]; s1 += r1[l][tx]; }
  d_beta[ch] = float(s0);
  d_gamma[ch] = float(s1);
}

// same thread mapping as bn_apply; mean == nullptr means "no normalisation" (plain ReLU backward)
__global__ void __launch_bounds__(256) bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                                                    int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                                    const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                                                    int relu, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16) {
  const int64_t n = live_rows(n_cap, n_dev);
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;
  const float inv_ppr = 1.f / float(ppr);
  const float inv_n = n > 0 ? 1.f / float(n) : 0.f;
  for (int64_t r0 = int64_t(blockIdx.x) * kApplyRows; r0 < n; r0 += int64_t(gridDim.x) * kApplyRows) {
    const int rows = int(n - r0 < kApplyRows ? n - r0 : kApplyRows);
    const float* xb = x + r0 * c;
    const float* dyb = dy + r0 * c;
    float* dxb = dx ? dx + r0 * c : nullptr;
    uint32_t* dx16b = dx16 ? reinterpret_cast<uint32_t*>(dx16 + r0 * c_pad) : nullptr;
    const int total = rows * ppr;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      int row = int((float(t) + 0.5f) * inv_ppr);
      int pr = t - row * ppr;
      if (pr < 0) { --row; pr += ppr; } else if (pr >= ppr) { ++row; pr -= ppr; }
      const int ch0 = pr << 1;
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ch = ch0 + e;
        if (ch < c) {
          float d = dyb[row * c + ch];
          float tv;
          if (mean) {
            const float is = invstd[ch], g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
            const float xh = (xb[row * c + ch] - mean[ch]) * is;
            if (relu && xh * g + b <= 0.f) d = 0.f;
            tv = g * is * (d - d_beta[ch] * inv_n - xh * d_gamma[ch] * inv_n);
          } else {
            tv = (relu && xb[row * c + ch] <= 0.f) ? 0.f : d;
          }
          if (dxb) dxb[row * c + ch] = tv;
          v[e] = tv;
        }
      }
      if (dx16b) dx16b[row * ppr + pr] = bn_pack_bf16x2(v[0], v[1]);
    }
  }
}

// ---- small problems: one launch per direction ----------------------------------------------------
// A block owns kCh channels for ALL live rows (threads (32, 16): 16 row lanes): statistics, running
// estimates and the normalise(+ReLU) pass in one kernel; the slab it re-reads stays in L1/L2.  Used
// when the row capacity is small (a whole C2 step is launch-latency bound).
constexpr int kSmallRows = 16384;
constexpr int kRowLanes = 16;

__device__ __forceinline__ float block_col_sum(float v, float (*red)[kCh], int tx, int ty) {
  red[ty][tx] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kRowLanes; ++i) t += red[i][tx];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kCh * kRowLanes) bn_fwd_small(
    const float* __restrict__ x, int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
    float* __restrict__ running_var, float momentum, float eps, int training, int relu, float* __restrict__ y,
    __nv_bfloat16* __restrict__ y16, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ float red[kRowLanes][kCh];
  const int64_t n = live_rows(n_cap, n_dev);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const bool on = ch < c;
  float mean, invstd;
  if (training) {
    float s = 0.f;
    if (on) for (int64_t r = ty; r < n; r += kRowLanes) s += x[r * c + ch];
    mean = n > 0 ? block_col_sum(s, red, tx, ty) / float(n) : 0.f;
    float m2 = 0.f;
    if (on) for (int64_t r = ty; r < n; r += kRowLanes) { const float d = x[r * c + ch] - mean; m2 += d * d; }
    m2 = block_col_sum(m2, red, tx, ty);
    invstd = n > 0 ? rsqrtf(m2 / float(n) + eps) : 0.f;
    if (on && ty == 0) {
      save_mean[ch] = mean;
      save_invstd[ch] = invstd;
      if (n > 0 && running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
      if (n > 1 && running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (m2 / float(n - 1));
    }
  } else {
    mean = on ? running_mean[ch] : 0.f;
    invstd = on ? rsqrtf(running_var[ch] + eps) : 0.f;
    if (on && ty == 0) { save_mean[ch] = mean; save_invstd[ch] = invstd; }
  }
  const int c_pad = (c + 7) & ~7;
  if (!on) {  // zero padding columns of the bf16 copy
    if (y16 && ch < c_pad)
      for (int64_t r = ty; r < n; r += kRowLanes) y16[r * c_pad + ch] = __float2bfloat16_rn(0.f);
    return;
  }
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  for (int64_t r = ty; r < n; r += kRowLanes) {
    float v = (x[r * c + ch] - mean) * invstd * g + b;
    v = (relu && v < 0.f) ? 0.f : v;
    if (y) y[r * c + ch] = v;
    if (y16) y16[r * c_pad + ch] = __float2bfloat16_rn(v);
  }
}

__global__ void __launch_bounds__(kCh * kRowLanes) bn_bwd_small(
    const float* __restrict__ x, const float* __restrict__ dy, int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean_,
    const float* __restrict__ invstd_, int relu, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16,
    float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  __shared__ float red[kRowLanes][kCh];
  const int64_t n = live_rows(n_cap, n_dev);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const bool on = ch < c;
  const float m = on ? mean_[ch] : 0.f, is = on ? invstd_[ch] : 0.f;
  const float g = (on && gamma) ? gamma[ch] : 1.f, b = (on && beta) ? beta[ch] : 0.f;
  float s0 = 0.f, s1 = 0.f;
  if (on)
    for (int64_t r = ty; r < n; r += kRowLanes) {
      const float xh = (x[r * c + ch] - m) * is;
      float d = dy[r * c + ch];
      if (relu && xh * g + b <= 0.f) d = 0.f;
      s0 += d;
      s1 += d * xh;
    }
  s0 = block_col_sum(s0, red, tx, ty);
  s1 = block_col_sum(s1, red, tx, ty);
  const int c_pad = (c + 7) & ~7;
  if (!on) {
    if (dx16 && ch < c_pad)
      for (int64_t r = ty; r < n; r += kRowLanes) dx16[r * c_pad + ch] = __float2bfloat16_rn(0.f);
    return;
  }
  if (ty == 0) { d_beta[ch] = s0; d_gamma[ch] = s1; }
  const float inv_n = n > 0 ? 1.f / float(n) : 0.f;
  for (int64_t r = ty; r < n; r += kRowLanes) {
    const float xh = (x[r * c + ch] - m) * is;
    float d = dy[r * c + ch];
    if (relu && xh * g + b <= 0.f) d = 0.f;
    const float t = g * is * (d - s0 * inv_n - xh * s1 * inv_n);
    if (dx) dx[r * c + ch] = t;
    if (dx16) dx16[r * c_pad + ch] = __float2bfloat16_rn(t);
  }
}

inline unsigned apply_blocks(int64_t rows) {
  int64_t b = ceil_div<int64_t>(rows > 0 ? rows : 1, kApplyRows);
  const int64_t cap = int64_t(sm_count()) * 16;
  return unsigned(b > cap ? cap : b);
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" size_t wfsp_bn_workspace_bytes(int64_t n_rows, int c) {
  return align_up(size_t(ceil_div<int64_t>(n_rows > 0 ? n_rows : 1, kRows)) * 2 * c * sizeof(float), 256);
}

extern "C" int wfsp_bn_relu_fwd_x(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                  int training, int relu, float* y, void* y_bf16, float* save_mean, float* save_invstd,
                                  void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad batch-norm sizes");
  WFSP_REQUIRE(y != nullptr || y_bf16 != nullptr, "batch norm needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* y16 = static_cast<__nv_bfloat16*>(y_bf16);
  if (!training) WFSP_REQUIRE(running_mean && running_var, "eval-mode batch norm needs running statistics");
  if (n_rows <= kSmallRows) {
    bn_fwd_small<<<ceil_div((c + 7) & ~7, kCh), dim3(kCh, kRowLanes), 0, st>>>(
        x, n_rows, n_rows_dev, c, gamma, beta, running_mean, running_var, momentum, eps, training, relu, y, y16,
        save_mean, save_invstd);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  if (training) {
    if (workspace == nullptr || workspace_bytes < wfsp_bn_workspace_bytes(n_rows, c))
      return set_error(WFSP_EWORKSPACE, "batch-norm workspace too small");
    float* part = static_cast<float*>(workspace);
    dim3 grid(unsigned(ceil_div<int64_t>(n_rows, kRows)), unsigned(ceil_div(c, kCh)));
    bn_partial_stats<<<grid, dim3(32, 8), 0, st>>>(x, n_rows, n_rows_dev, c, part);
    bn_finalize_stats<<<ceil_div(c, kCh), dim3(kCh, kFinLanes), 0, st>>>(part, n_rows, n_rows_dev, c, eps, momentum, running_mean,
                                                        running_var, save_mean, save_invstd);
    count_launches(2);
  } else {
    bn_eval_stats<<<ceil_div(c, 128), 128, 0, st>>>(c, eps, running_mean, running_var, save_mean, save_invstd);
    count_launches(1);
  }
  bn_apply<<<apply_blocks(n_rows), 256, 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, save_mean,
                                                                   save_invstd, relu, y, y16);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                int training, int relu, float* y, float* save_mean, float* save_invstd,
                                void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_bn_relu_fwd_x(x, n_rows, n_rows_dev, c, gamma, beta, running_mean, running_var, momentum, eps, training,
                            relu, y, nullptr, save_mean, save_invstd, workspace, workspace_bytes, stream);
}

extern "C" int wfsp_bn_relu_bwd_x(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c,
                                  const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_invstd, int relu, float* dx, void* dx_bf16, float* d_gamma,
                                  float* d_beta, void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad batch-norm sizes");
  WFSP_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "batch norm backward needs at least one output");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* dx16 = static_cast<__nv_bfloat16*>(dx_bf16);
  if (n_rows == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_gamma, 0, size_t(c) * 4, st));
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_beta, 0, size_t(c) * 4, st));
    return WFSP_OK;
  }
  if (n_rows <= kSmallRows) {
    bn_bwd_small<<<ceil_div((c + 7) & ~7, kCh), dim3(kCh, kRowLanes), 0, st>>>(
        x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, dx, dx16, d_gamma, d_beta);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  if (workspace == nullptr || workspace_bytes < wfsp_bn_workspace_bytes(n_rows, c))
    return set_error(WFSP_EWORKSPACE, "batch-norm workspace too small");
  float* part = static_cast<float*>(workspace);
  dim3 grid(unsigned(ceil_div<int64_t>(n_rows, kRows)), unsigned(ceil_div(c, kCh)));
  bn_bwd_partial<<<grid, dim3(32, 8), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, part);
  bn_bwd_finalize<<<ceil_div(c, kCh), dim3(kCh, kFinLanes), 0, st>>>(part, n_rows, n_rows_dev, c, d_gamma, d_beta);
  bn_bwd_apply<<<apply_blocks(n_rows), 256, 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean,
                                                                       save_invstd, d_gamma, d_beta, relu, dx, dx16);
  count_launches(3);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c,
                                const float* gamma, const float* beta, const float* save_mean,
                                const float* save_invstd, int relu, float* dx, float* d_gamma, float* d_beta,
                                void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_bn_relu_bwd_x(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, dx, nullptr, d_gamma,
                            d_beta, workspace, workspace_bytes, stream);
}

// Activation-only blocks (a convolution followed by ReLU / nothing, no BatchNorm): the same streaming
// kernels with the normalisation switched off.
extern "C" int wfsp_act_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, int relu, float* y,
                            void* y_bf16, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad sizes");
  WFSP_REQUIRE(y != nullptr || y_bf16 != nullptr, "needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  bn_apply<<<apply_blocks(n_rows), 256, 0, as_stream(stream)>>>(
      x, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, relu, y, static_cast<__nv_bfloat16*>(y_bf16));
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_act_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c, int relu,
                            float* dx, void* dx_bf16, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad sizes");
  WFSP_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  bn_bwd_apply<<<apply_blocks(n_rows), 256, 0, as_stream(stream)>>>(
      x, dy, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, relu, dx,
      static_cast<__nv_bfloat16*>(dx_bf16));
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
