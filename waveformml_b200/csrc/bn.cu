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

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace wfsp {
namespace {

constexpr int kRows = 1024;  // rows per partial (forward statistics of the per-layer path)
constexpr int kRowsBwd = 256;  // rows per partial of the backward sums: more, smaller CTAs keep enough loads in flight
constexpr int kCh = 32;      // channels per block (one 128-byte row segment)
constexpr int kPartLanes = 32;  // row lanes of the partial-statistics kernels: 1024 threads per block keep enough loads in flight

__device__ __forceinline__ int64_t live_rows(int64_t n, const int32_t* n_dev) { return n_dev ? int64_t(*n_dev) : n; }

// grid (ceil(cap/kRows), ceil(c/kCh)), block (32, 8).  One pass over x: every thread accumulates sum and
// sum of squares of (x - K), K = its first value (a shift that removes the cancellation of the textbook
// formula), turns them into (count, mean, M2) and the eight row lanes are merged with Chan's formula in a
// fixed order.
__global__ void __launch_bounds__(kCh * kPartLanes) bn_partial_stats(const float* __restrict__ x, int64_t n_cap,
                                                        const int32_t* __restrict__ n_dev, int c,
                                                        float* __restrict__ part /* [nblk][2][c] mean, M2 */) {
  __shared__ float s_cnt[kPartLanes][kCh], s_mean[kPartLanes][kCh], s_m2[kPartLanes][kCh];
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t r0 = int64_t(blockIdx.x) * kRows;
  if (r0 >= n) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.y * kCh + tx;
  const int rows = int(n - r0 < kRows ? n - r0 : kRows);
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  if (ch < c && ty < rows) {
    const float* xp = x + r0 * c + ch;
    const float K = xp[int64_t(ty) * c];
    float sd = 0.f, sq = 0.f;
#pragma unroll 8
    for (int r = ty; r < rows; r += kPartLanes) {
      const float d = xp[int64_t(r) * c] - K;
      sd += d;
      sq += d * d;
    }
    cnt = float((rows - ty + kPartLanes - 1) / kPartLanes);
    mean = K + sd / cnt;
    m2 = sq - sd * sd / cnt;
    if (m2 < 0.f) m2 = 0.f;
  }
  s_cnt[ty][tx] = cnt; s_mean[ty][tx] = mean; s_m2[ty][tx] = m2;
  __syncthreads();
  if (ty == 0 && ch < c) {
    for (int l = 1; l < kPartLanes; ++l) {
      const float nb = s_cnt[l][tx];
      if (nb > 0.f) {
        const float delta = s_mean[l][tx] - mean, tot = cnt + nb;
        mean += delta * nb / tot;
        m2 += s_m2[l][tx] + delta * delta * cnt * nb / tot;
        cnt = tot;
      }
    }
    part[(int64_t(blockIdx.x) * 2 + 0) * c + ch] = mean;
    part[(int64_t(blockIdx.x) * 2 + 1) * c + ch] = m2;
  }
}

// One block per 32 channels; 32 row lanes each fold every 32nd partial, then lane 0 adds the 32 lane results
// in lane order (a fixed order, so the statistics do not depend on scheduling).  Partials (n_b, mean_b, M2_b)
// are folded as plain double-precision sums N = sum n_b, A = sum n_b mean_b, Q = sum (M2_b + n_b mean_b^2),
// and mean = A / N, M2 = Q - N mean^2: no division per partial (FP64 division is slow on this part), and in
// double the subtraction costs ~1e-16 * mean^2 / var of relative accuracy.
constexpr int kFinLanes = 32;
constexpr int kFinSegMax = 32;  // the partial list is cut into up to this many segments, one block each per channel group

// grid (channel groups, S).  Block (g, seg) folds segment seg of the partial list; the last of the S blocks of
// a channel group to finish (ticket counter) adds the S segment results in segment order and writes the
// statistics.  `inter` = [S][2][c] doubles, `tickets` = [channel groups] zeroed ints (left zero again).
__global__ void __launch_bounds__(kCh * kFinLanes) bn_finalize_stats(const float* __restrict__ part, int chunk_rows, int64_t n_cap,
                                                         const int32_t* __restrict__ n_dev, int c, float eps,
                                                         float momentum, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var,
                                                         float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                                         double* __restrict__ inter, unsigned* __restrict__ tickets) {
  __shared__ double s_a[kFinLanes][kCh], s_q[kFinLanes][kCh];
  __shared__ unsigned s_ticket;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const int S = gridDim.y, seg = blockIdx.y;
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t nblk = n > 0 ? (n + chunk_rows - 1) / chunk_rows : 0;
  const int64_t per = (nblk + S - 1) / S;
  const int64_t lo = seg * per, hi = lo + per < nblk ? lo + per : nblk;
  double a = 0.0, q = 0.0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t b = lo + ty; b < hi; b += kFinLanes) {
      const double nb = double(b + 1 < nblk ? chunk_rows : n - b * chunk_rows);
      const double mb = part[(b * 2 + 0) * c + ch], m2b = part[(b * 2 + 1) * c + ch];
      a += nb * mb;
      q += m2b + nb * mb * mb;
    }
  }
  s_a[ty][tx] = a; s_q[ty][tx] = q;
  __syncthreads();
  if (ty == 0) {
    for (int l = 1; l < kFinLanes; ++l) { a += s_a[l][tx]; q += s_q[l][tx]; }
  }
  if (S > 1) {
    if (ty == 0 && ch < c) {
      inter[(int64_t(seg) * 2 + 0) * c + ch] = a;
      inter[(int64_t(seg) * 2 + 1) * c + ch] = q;
    }
    __threadfence();
    __syncthreads();
    if (tx == 0 && ty == 0) s_ticket = atomicAdd(&tickets[blockIdx.x], 1u);
    __syncthreads();
    if (s_ticket != unsigned(S - 1)) return;  // not the last block of this channel group
    __threadfence();
    if (tx == 0 && ty == 0) tickets[blockIdx.x] = 0u;
    if (ty == 0 && ch < c) {
      a = 0.0; q = 0.0;
      for (int sg = 0; sg < S; ++sg) {
        a += __ldcg(&inter[(int64_t(sg) * 2 + 0) * c + ch]);
        q += __ldcg(&inter[(int64_t(sg) * 2 + 1) * c + ch]);
      }
    }
  }
  if (ty != 0 || ch >= c) return;
  if (n <= 0) { save_mean[ch] = 0.f; save_invstd[ch] = 0.f; return; }
  const double cnt = double(n), mean = a / cnt;
  double m2 = q - cnt * mean * mean;
  if (m2 < 0.0) m2 = 0.0;
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

// Streaming normalise(+ReLU) pass.  Block (bdx, bdy): thread x owns one pair of adjacent channels and keeps
// its statistics / affine parameters in registers, thread y strides over the kApplyRows rows of the block's
// chunk; a warp reads 256 contiguous bytes of a row and writes 256 (fp32) / 128 (bf16).  Outputs (each
// optional): y fp32 [rows, c] and y16 bf16 [rows, c_pad] (c_pad = c rounded up to 8, padding written as
// zero) -- the operand format of the tcgen05 convolution kernels, so the next layer needs no cast pass.
// mean == nullptr switches the normalisation off (plain ReLU / cast).
constexpr int kApplyRows = 64;  // rows per CTA pass at most (large inputs); small inputs take fewer, see apply_rows()

// VEC2: c is even and the buffers are 8-byte aligned, so a channel pair is one float2.  Four rows are
// loaded before any is used (independent loads in flight: these kernels are pure HBM streams).
constexpr int kApplyUnroll = 4;

// Statistics of the fused small path: every CTA folds the (few) per-chunk partials of its own channels itself
// -- the same sums, in the same order, as bn_finalize_stats, so all CTAs get bit-identical values -- and CTA 0
// also records them (save_mean / save_invstd / running estimates).  One launch, no cluster barrier.
struct PartStats {
  const float* part;  // [chunks][2][c] (mean, M2) per chunk of chunk_rows rows, or nullptr
  int chunk_rows;
  float eps, momentum;
  float* running_mean; float* running_var; float* save_mean; float* save_invstd;
};

__device__ __forceinline__ void stats_from_partials(const PartStats& ps, int64_t n, int c, int ch, float& mean, float& invstd,
                                                    bool record) {
  const int64_t nblk = (n + ps.chunk_rows - 1) / ps.chunk_rows;
  double a = 0.0, q = 0.0;
  for (int64_t b = 0; b < nblk; ++b) {
    const double nb = double(b + 1 < nblk ? ps.chunk_rows : n - b * ps.chunk_rows);
    const double mb = ps.part[(b * 2 + 0) * c + ch], m2b = ps.part[(b * 2 + 1) * c + ch];
    a += nb * mb;
    q += m2b + nb * mb * mb;
  }
  const double cnt = double(n), mu = a / cnt;
  double m2 = q - cnt * mu * mu;
  if (m2 < 0.0) m2 = 0.0;
  mean = float(mu);
  invstd = float(1.0 / sqrt(m2 / cnt + double(ps.eps)));
  if (record) {
    ps.save_mean[ch] = mean;
    ps.save_invstd[ch] = invstd;
    if (ps.running_mean) ps.running_mean[ch] = (1.f - ps.momentum) * ps.running_mean[ch] + ps.momentum * mean;
    if (ps.running_var && cnt > 1.0)
      ps.running_var[ch] = (1.f - ps.momentum) * ps.running_var[ch] + ps.momentum * float(m2 / (cnt - 1.0));
  }
}

// FOLD: the fused small path (statistics folded from the partials by every CTA); the streaming variant is kept lean
// (<= 64 registers, four CTAs per SM) -- these kernels are bound by the bytes they keep in flight.
template <bool VEC2, bool FOLD>
__global__ void __launch_bounds__(256, FOLD ? 1 : 4) bn_apply(const float* __restrict__ x, int64_t n_cap,
                                                const int32_t* __restrict__ n_dev, int c,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                int relu, float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                                                const PartStats ps, int rows_per_cta, const DropSpec drop) {
  const int64_t n = live_rows(n_cap, n_dev);
  const unsigned long long dkey = drop.p > 0.f ? drop_key(drop) : 0ull;
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;  // channel pairs per (padded) row
  uint32_t* y16w = reinterpret_cast<uint32_t*>(y16);
  if (int64_t(blockIdx.x) * rows_per_cta >= n && !(FOLD && blockIdx.x == 0)) return;
  __shared__ float s_m[FOLD ? 512 : 1], s_is[FOLD ? 512 : 1];  // fused small path: statistics of this CTA's channels (c <= 512)
  if (FOLD && n > 0) {
    for (int ch = threadIdx.y * blockDim.x + threadIdx.x; ch < c; ch += blockDim.x * blockDim.y)
      stats_from_partials(ps, n, c, ch, s_m[ch], s_is[ch], blockIdx.x == 0);
    __syncthreads();
  }
  for (int pc = threadIdx.x; pc < ppr; pc += blockDim.x) {
    const int ch = pc << 1;
    float m[2] = {0.f, 0.f}, is[2] = {1.f, 1.f}, g[2] = {1.f, 1.f}, b[2] = {0.f, 0.f};
    const bool on[2] = {ch < c, ch + 1 < c};
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if ((mean || FOLD) && on[e]) {
        if (FOLD) { m[e] = s_m[ch + e]; is[e] = s_is[ch + e]; }
        else { m[e] = mean[ch + e]; is[e] = invstd[ch + e]; }
        if (gamma) g[e] = gamma[ch + e];
        if (beta) b[e] = beta[ch + e];
      }
    for (int64_t r0 = int64_t(blockIdx.x) * rows_per_cta; r0 < n; r0 += int64_t(gridDim.x) * rows_per_cta) {
      const int64_t r_end = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
      for (int64_t rb = r0 + threadIdx.y; rb < r_end; rb += int64_t(kApplyUnroll) * blockDim.y) {
        float xv[kApplyUnroll][2];
#pragma unroll
        for (int u = 0; u < kApplyUnroll; ++u) {
          const int64_t r = rb + int64_t(u) * blockDim.y;
          xv[u][0] = xv[u][1] = 0.f;
          if (r < r_end) {
            if (VEC2) {
              if (on[0]) { const float2 t = *reinterpret_cast<const float2*>(x + r * c + ch); xv[u][0] = t.x; xv[u][1] = t.y; }
            } else {
              if (on[0]) xv[u][0] = x[r * c + ch];
              if (on[1]) xv[u][1] = x[r * c + ch + 1];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kApplyUnroll; ++u) {
          const int64_t r = rb + int64_t(u) * blockDim.y;
          if (r >= r_end) continue;
          float v[2] = {0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (on[e]) {
              float tv = xv[u][e];
              if (mean || FOLD) tv = (tv - m[e]) * is[e] * g[e] + b[e];
              if (relu && tv < 0.f) tv = 0.f;
              if (drop.p > 0.f) tv *= drop_factor(dkey, drop.p, drop.scale, (unsigned long long)(r * c + ch + e));
              v[e] = tv;
            }
          }
          if (y) {
            if (VEC2) { if (on[0]) *reinterpret_cast<float2*>(y + r * c + ch) = make_float2(v[0], v[1]); }
            else { if (on[0]) y[r * c + ch] = v[0]; if (on[1]) y[r * c + ch + 1] = v[1]; }
          }
          if (y16w) y16w[r * ppr + pc] = bn_pack_bf16x2(v[0], v[1]);
        }
      }
    }
  }
}

// backward partials: sum(dy') and sum(dy' * xhat) per (row chunk, channel); dy' = dy masked by relu.
// Same thread mapping as the apply kernels: thread x owns one pair of adjacent channels (statistics and affine
// parameters in registers, float2 accesses when VEC2), thread y strides over the chunk's rows with four rows in
// flight; a warp reads 256 contiguous bytes per row.  grid = row chunks of kRowsBwd, block = apply_block(c).
template <bool VEC2>
__global__ void __launch_bounds__(256, 4) bn_bwd_partial(const float* __restrict__ x, const float* __restrict__ dy,
                                                      int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                      int relu, float* __restrict__ part /* [nblk][2][c] */,
                                                      const DropSpec drop) {
  __shared__ float red[2][2][256];  // [sum kind][channel of the pair][thread]
  const int64_t n = live_rows(n_cap, n_dev);
  const unsigned long long dkey = drop.p > 0.f ? drop_key(drop) : 0ull;
  const int64_t r0 = int64_t(blockIdx.x) * kRowsBwd;
  if (r0 >= n) return;
  const int64_t r_end = r0 + kRowsBwd < n ? r0 + kRowsBwd : n;
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;
  const int tflat = threadIdx.y * blockDim.x + threadIdx.x;
  for (int pc0 = 0; pc0 < ppr; pc0 += blockDim.x) {
    const int pc = pc0 + threadIdx.x, ch = pc << 1;
    const bool on[2] = {pc < ppr && ch < c, pc < ppr && ch + 1 < c};
    float m[2] = {0.f, 0.f}, is[2] = {0.f, 0.f}, g[2] = {1.f, 1.f}, b[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (on[e]) {
        m[e] = mean[ch + e]; is[e] = invstd[ch + e];
        if (gamma) g[e] = gamma[ch + e];
        if (beta) b[e] = beta[ch + e];
      }
    float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
    for (int64_t rb = r0 + threadIdx.y; rb < r_end; rb += int64_t(kApplyUnroll) * blockDim.y) {
      float xv[kApplyUnroll][2], dv[kApplyUnroll][2];
#pragma unroll
      for (int u = 0; u < kApplyUnroll; ++u) {
        const int64_t r = rb + int64_t(u) * blockDim.y;
        xv[u][0] = xv[u][1] = dv[u][0] = dv[u][1] = 0.f;
        if (r < r_end) {
          if (VEC2) {
            if (on[0]) {
              const float2 t = *reinterpret_cast<const float2*>(x + r * c + ch);
              const float2 d = *reinterpret_cast<const float2*>(dy + r * c + ch);
              xv[u][0] = t.x; xv[u][1] = t.y; dv[u][0] = d.x; dv[u][1] = d.y;
            }
          } else {
            if (on[0]) { xv[u][0] = x[r * c + ch]; dv[u][0] = dy[r * c + ch]; }
            if (on[1]) { xv[u][1] = x[r * c + ch + 1]; dv[u][1] = dy[r * c + ch + 1]; }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kApplyUnroll; ++u) {
        if (rb + int64_t(u) * blockDim.y >= r_end) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float xh = (xv[u][e] - m[e]) * is[e];
          float d = dv[u][e];
          if (drop.p > 0.f && on[e])
            d *= drop_factor(dkey, drop.p, drop.scale, (unsigned long long)((rb + int64_t(u) * blockDim.y) * c + ch + e));
          if (relu && xh * g[e] + b[e] <= 0.f) d = 0.f;
          s0[e] += d;
          s1[e] += d * xh;
        }
      }
    }
    // combine the row lanes (threadIdx.y) of every channel pair in a fixed order
    red[0][0][tflat] = s0[0]; red[0][1][tflat] = s0[1]; red[1][0][tflat] = s1[0]; red[1][1][tflat] = s1[1];
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (!on[e]) continue;
        float t0 = 0.f, t1 = 0.f;
        for (int l = 0; l < int(blockDim.y); ++l) {
          t0 += red[0][e][l * blockDim.x + threadIdx.x];
          t1 += red[1][e][l * blockDim.x + threadIdx.x];
        }
        part[(int64_t(blockIdx.x) * 2 + 0) * c + ch + e] = t0;
        part[(int64_t(blockIdx.x) * 2 + 1) * c + ch + e] = t1;
      }
    }
    __syncthreads();
  }
}

// grid (channel groups, S): block (g, seg) folds segment seg of the partial list (chunks of `chunk_rows` rows); with
// S > 1 the last block of a channel group to finish (ticket) adds the S segment results in segment order -- the
// same scheme as bn_finalize_stats, so the sums do not depend on scheduling.
__global__ void __launch_bounds__(kCh * kFinLanes) bn_bwd_finalize(const float* __restrict__ part, int chunk_rows, int64_t n_cap,
                                                       const int32_t* __restrict__ n_dev, int c,
                                                       float* __restrict__ d_gamma, float* __restrict__ d_beta,
                                                       double* __restrict__ inter, unsigned* __restrict__ tickets,
                                                       int nblk_fixed /* >= 0: length of the partial list (one entry per CTA
                                                                         of the streaming kernel), else derived from the rows */) {
  __shared__ double r0[kFinLanes][kCh], r1[kFinLanes][kCh];
  __shared__ unsigned s_ticket;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const int S = gridDim.y, seg = blockIdx.y;
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t nblk = nblk_fixed >= 0 ? int64_t(nblk_fixed) : (n > 0 ? (n + chunk_rows - 1) / chunk_rows : 0);
  const int64_t per = (nblk + S - 1) / S;
  const int64_t lo = seg * per, hi = lo + per < nblk ? lo + per : nblk;
  double s0 = 0.0, s1 = 0.0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t b = lo + ty; b < hi; b += kFinLanes) {
      s0 += part[(b * 2 + 0) * c + ch];
      s1 += part[(b * 2 + 1) * c + ch];
    }
  }
  r0[ty][tx] = s0; r1[ty][tx] = s1;
  __syncthreads();
  if (ty == 0)
    for (int l = 1; l < kFinLanes; ++l) { s0 += r0[l][tx]; s1 += r1[l][tx]; }
  if (S > 1) {
    if (ty == 0 && ch < c) {
      inter[(int64_t(seg) * 2 + 0) * c + ch] = s0;
      inter[(int64_t(seg) * 2 + 1) * c + ch] = s1;
    }
    __threadfence();
    __syncthreads();
    if (tx == 0 && ty == 0) s_ticket = atomicAdd(&tickets[blockIdx.x], 1u);
    __syncthreads();
    if (s_ticket != unsigned(S - 1)) return;
    __threadfence();
    if (tx == 0 && ty == 0) tickets[blockIdx.x] = 0u;
    if (ty == 0 && ch < c) {
      s0 = 0.0; s1 = 0.0;
      for (int sg = 0; sg < S; ++sg) {
        s0 += __ldcg(&inter[(int64_t(sg) * 2 + 0) * c + ch]);
        s1 += __ldcg(&inter[(int64_t(sg) * 2 + 1) * c + ch]);
      }
    }
  }
  if (ty != 0 || ch >= c) return;
  d_beta[ch] = float(s0);
  d_gamma[ch] = float(s1);
}

// same thread mapping as bn_apply; mean == nullptr means "no normalisation" (plain ReLU backward)
template <bool VEC2, bool FOLD>
__global__ void __launch_bounds__(256, FOLD ? 1 : 4) bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                                                    int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                                    const float* __restrict__ d_gamma, const float* __restrict__ d_beta,
                                                    int relu, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16,
                                                    int rows_per_cta, const float* __restrict__ part, int chunk_rows,
                                                    float* __restrict__ d_gamma_out, float* __restrict__ d_beta_out,
                                                    const DropSpec drop) {
  const int64_t n = live_rows(n_cap, n_dev);
  const unsigned long long dkey = drop.p > 0.f ? drop_key(drop) : 0ull;
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;
  const float inv_n = n > 0 ? 1.f / float(n) : 0.f;
  uint32_t* dx16w = reinterpret_cast<uint32_t*>(dx16);
  // fused small path: the (few) per-chunk sums of the dgrad epilogue are folded by every CTA for its own use --
  // same order everywhere, so all CTAs hold bit-identical sums -- and CTA 0 records d_gamma / d_beta (c <= 512)
  __shared__ float s_db[FOLD ? 512 : 1], s_dg[FOLD ? 512 : 1];
  if (FOLD) {
    const int64_t nblk = n > 0 ? (n + chunk_rows - 1) / chunk_rows : 0;
    for (int ch = threadIdx.y * blockDim.x + threadIdx.x; ch < c; ch += blockDim.x * blockDim.y) {
      double a = 0.0, q = 0.0;
      for (int64_t b = 0; b < nblk; ++b) {
        a += part[(b * 2 + 0) * c + ch];
        q += part[(b * 2 + 1) * c + ch];
      }
      s_db[ch] = float(a); s_dg[ch] = float(q);
      if (blockIdx.x == 0) { d_beta_out[ch] = float(a); d_gamma_out[ch] = float(q); }
    }
    __syncthreads();
    d_beta = s_db; d_gamma = s_dg;
  }
  for (int pc = threadIdx.x; pc < ppr; pc += blockDim.x) {
    const int ch = pc << 1;
    float m[2] = {0.f, 0.f}, is[2] = {1.f, 1.f}, g[2] = {1.f, 1.f}, b[2] = {0.f, 0.f}, db[2] = {0.f, 0.f}, dg[2] = {0.f, 0.f};
    const bool on[2] = {ch < c, ch + 1 < c};
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (mean && on[e]) {
        m[e] = mean[ch + e]; is[e] = invstd[ch + e];
        if (gamma) g[e] = gamma[ch + e];
        if (beta) b[e] = beta[ch + e];
        db[e] = d_beta[ch + e] * inv_n; dg[e] = d_gamma[ch + e] * inv_n;
      }
    for (int64_t r0 = int64_t(blockIdx.x) * rows_per_cta; r0 < n; r0 += int64_t(gridDim.x) * rows_per_cta) {
      const int64_t r_end = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
      for (int64_t rb = r0 + threadIdx.y; rb < r_end; rb += int64_t(kApplyUnroll) * blockDim.y) {
        float xv[kApplyUnroll][2], dv[kApplyUnroll][2];
#pragma unroll
        for (int u = 0; u < kApplyUnroll; ++u) {
          const int64_t r = rb + int64_t(u) * blockDim.y;
          xv[u][0] = xv[u][1] = dv[u][0] = dv[u][1] = 0.f;
          if (r < r_end) {
            if (VEC2) {
              if (on[0]) {
                const float2 t = *reinterpret_cast<const float2*>(x + r * c + ch);
                const float2 d = *reinterpret_cast<const float2*>(dy + r * c + ch);
                xv[u][0] = t.x; xv[u][1] = t.y; dv[u][0] = d.x; dv[u][1] = d.y;
              }
            } else {
              if (on[0]) { xv[u][0] = x[r * c + ch]; dv[u][0] = dy[r * c + ch]; }
              if (on[1]) { xv[u][1] = x[r * c + ch + 1]; dv[u][1] = dy[r * c + ch + 1]; }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kApplyUnroll; ++u) {
          const int64_t r = rb + int64_t(u) * blockDim.y;
          if (r >= r_end) continue;
          float v[2] = {0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (on[e]) {
              float d = dv[u][e];
              if (drop.p > 0.f) d *= drop_factor(dkey, drop.p, drop.scale, (unsigned long long)(r * c + ch + e));
              if (mean) {
                const float xh = (xv[u][e] - m[e]) * is[e];
                if (relu && xh * g[e] + b[e] <= 0.f) d = 0.f;
                v[e] = g[e] * is[e] * (d - db[e] - xh * dg[e]);
              } else {
                v[e] = (relu && xv[u][e] <= 0.f) ? 0.f : d;
              }
            }
          }
          if (dx) {
            if (VEC2) { if (on[0]) *reinterpret_cast<float2*>(dx + r * c + ch) = make_float2(v[0], v[1]); }
            else { if (on[0]) dx[r * c + ch] = v[0]; if (on[1]) dx[r * c + ch + 1] = v[1]; }
          }
          if (dx16w) dx16w[r * ppr + pc] = bn_pack_bf16x2(v[0], v[1]);
        }
      }
    }
  }
}

// ---- small problems: one launch per direction ----------------------------------------------------
// A block owns kCh channels for ALL live rows (threads (32, 16): 16 row lanes): statistics, running
// estimates and the normalise(+ReLU) pass in one kernel; the slab it re-reads stays in L1/L2.  Used
// when the row capacity is small (a whole C2 step is launch-latency bound).
constexpr int kSmallRows = 16384;
constexpr int kFoldRows = 2048;     // (expected live) rows up to which the apply kernel folds the conv partials itself
constexpr int kClusterRows = 4096;  // (expected live) rows up to which the backward runs as one cluster launch
constexpr int kRowLanes = 16;
constexpr int kClusterY = 8;  // CTAs of one thread-block cluster: they split the rows of a channel group

// Column sums over ALL rows of the cluster: per-CTA partial in shared memory, cluster barrier, then every CTA
// adds the kClusterY partials in rank order through distributed shared memory (same order everywhere, so
// all CTAs hold bit-identical statistics).  NV values per channel are reduced at once.
template <int NV>
__device__ __forceinline__ void cluster_col_sum(float (&v)[NV], float (*red)[kCh], float (*part)[kCh], int tx, int ty,
                                                cg::cluster_group& cluster) {
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    red[ty][tx] = v[q];
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < kRowLanes; ++i) t += red[i][tx];
      part[q][tx] = t;
    }
    __syncthreads();
  }
  cluster.sync();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    float t = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += cluster.map_shared_rank(&part[q][0], r)[tx];
    v[q] = t;
  }
  cluster.sync();  // the partials may be overwritten (or the CTA may exit) only after every peer has read them
}

// grid (channel groups, kClusterY), cluster (1, kClusterY, 1), block (32, 16)
__global__ void __launch_bounds__(kCh * kRowLanes) bn_fwd_small(
    const float* __restrict__ x, int64_t n_cap, const int32_t* __restrict__ n_dev, int c,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
    float* __restrict__ running_var, float momentum, float eps, int training, int relu, float* __restrict__ y,
    __nv_bfloat16* __restrict__ y16, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ float red[kRowLanes][kCh], red2[kRowLanes][kCh], red3[kRowLanes][kCh];
  __shared__ float part[3][kCh];
  cg::cluster_group cluster = cg::this_cluster();
  const int64_t n = live_rows(n_cap, n_dev);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const bool on = ch < c;
  const int64_t row0 = int64_t(cluster.block_rank()) * kRowLanes + ty, rstep = int64_t(cluster.num_blocks()) * kRowLanes;
  const bool lead = cluster.block_rank() == 0 && ty == 0;
  float mean, invstd;
  if (training) {
    // one pass: per-thread (count, mean, M2) from sums shifted by the thread's first value, merged with Chan's
    // formula over the 16 row lanes (by lane 0, in lane order) and then over the cluster's CTAs in rank order
    float cnt = 0.f, mu = 0.f, m2 = 0.f;
    if (on && row0 < n) {
      const float K = x[row0 * c + ch];
      float sd = 0.f, sq = 0.f;
      for (int64_t r = row0; r < n; r += rstep) {
        const float d = x[r * c + ch] - K;
        sd += d;
        sq += d * d;
        cnt += 1.f;
      }
      mu = K + sd / cnt;
      m2 = fmaxf(sq - sd * sd / cnt, 0.f);
    }
    red[ty][tx] = cnt; red2[ty][tx] = mu; red3[ty][tx] = m2;
    __syncthreads();
    if (ty == 0) {
      for (int l = 1; l < kRowLanes; ++l) {
        const float nb = red[l][tx];
        if (nb > 0.f) {
          const float delta = red2[l][tx] - mu, tot = cnt + nb;
          mu += delta * nb / tot;
          m2 += red3[l][tx] + delta * delta * cnt * nb / tot;
          cnt = tot;
        }
      }
      part[0][tx] = cnt; part[1][tx] = mu; part[2][tx] = m2;
    }
    cluster.sync();
    cnt = 0.f; mu = 0.f; m2 = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
      const float nb = cluster.map_shared_rank(&part[0][0], r)[tx];
      if (nb > 0.f) {
        const float delta = cluster.map_shared_rank(&part[1][0], r)[tx] - mu, tot = cnt + nb;
        mu += delta * nb / tot;
        m2 += cluster.map_shared_rank(&part[2][0], r)[tx] + delta * delta * cnt * nb / tot;
        cnt = tot;
      }
    }
    cluster.sync();  // peers have read this CTA's partials
    mean = mu;
    invstd = n > 0 ? rsqrtf(m2 / float(n) + eps) : 0.f;
    if (on && lead) {
      save_mean[ch] = mean;
      save_invstd[ch] = invstd;
      if (n > 0 && running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
      if (n > 1 && running_var) running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (m2 / float(n - 1));
    }
  } else {
    mean = on ? running_mean[ch] : 0.f;
    invstd = on ? rsqrtf(running_var[ch] + eps) : 0.f;
    if (on && lead) { save_mean[ch] = mean; save_invstd[ch] = invstd; }
  }
  const int c_pad = (c + 7) & ~7;
  if (!on) {  // zero padding columns of the bf16 copy
    if (y16 && ch < c_pad)
      for (int64_t r = row0; r < n; r += rstep) y16[r * c_pad + ch] = __float2bfloat16_rn(0.f);
    return;
  }
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  for (int64_t r = row0; r < n; r += rstep) {
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
    float* __restrict__ d_gamma, float* __restrict__ d_beta, const DropSpec drop) {
  __shared__ float red[kRowLanes][kCh];
  __shared__ float part[2][kCh];
  const unsigned long long dkey = drop.p > 0.f ? drop_key(drop) : 0ull;
  cg::cluster_group cluster = cg::this_cluster();
  const int64_t n = live_rows(n_cap, n_dev);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * kCh + tx;
  const bool on = ch < c;
  const int64_t row0 = int64_t(cluster.block_rank()) * kRowLanes + ty, rstep = int64_t(cluster.num_blocks()) * kRowLanes;
  const float m = on ? mean_[ch] : 0.f, is = on ? invstd_[ch] : 0.f;
  const float g = (on && gamma) ? gamma[ch] : 1.f, b = (on && beta) ? beta[ch] : 0.f;
  float s[2] = {0.f, 0.f};
  if (on)
    for (int64_t r = row0; r < n; r += rstep) {
      const float xh = (x[r * c + ch] - m) * is;
      float d = dy[r * c + ch];
      if (drop.p > 0.f) d *= drop_factor(dkey, drop.p, drop.scale, (unsigned long long)(r * c + ch));
      if (relu && xh * g + b <= 0.f) d = 0.f;
      s[0] += d;
      s[1] += d * xh;
    }
  cluster_col_sum<2>(s, red, part, tx, ty, cluster);
  const float s0 = s[0], s1 = s[1];
  const int c_pad = (c + 7) & ~7;
  if (!on) {
    if (dx16 && ch < c_pad)
      for (int64_t r = row0; r < n; r += rstep) dx16[r * c_pad + ch] = __float2bfloat16_rn(0.f);
    return;
  }
  if (cluster.block_rank() == 0 && ty == 0) { d_beta[ch] = s0; d_gamma[ch] = s1; }
  const float inv_n = n > 0 ? 1.f / float(n) : 0.f;
  for (int64_t r = row0; r < n; r += rstep) {
    const float xh = (x[r * c + ch] - m) * is;
    float d = dy[r * c + ch];
    if (drop.p > 0.f) d *= drop_factor(dkey, drop.p, drop.scale, (unsigned long long)(r * c + ch));
    if (relu && xh * g + b <= 0.f) d = 0.f;
    const float t = g * is * (d - s0 * inv_n - xh * s1 * inv_n);
    if (dx) dx[r * c + ch] = t;
    if (dx16) dx16[r * c_pad + ch] = __float2bfloat16_rn(t);
  }
}

// launch of the small kernels as thread-block clusters of kClusterY CTAs along y
template <typename... KArgs, typename... Args>
cudaError_t launch_cluster(void (*kernel)(KArgs...), int groups, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(groups), kClusterY, 1);
  cfg.blockDim = dim3(kCh, kRowLanes, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = kClusterY;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// scratch of bn_finalize_stats behind a partial list of `chunks` entries: [S][2][c] doubles + tickets
inline int fin_segments(int64_t chunks) {
  int64_t s = chunks / 128;
  return int(s < 1 ? 1 : (s > kFinSegMax ? kFinSegMax : s));
}
inline size_t fin_scratch_bytes(int c) { return size_t(kFinSegMax) * 2 * c * sizeof(double) + 1024; }

int launch_finalize(const float* part, int chunk_rows, int64_t n_rows, const int32_t* n_rows_dev, int c, float eps,
                    float momentum, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                    void* scratch, cudaStream_t st) {
  const int S = fin_segments(ceil_div<int64_t>(n_rows, chunk_rows));
  double* inter = static_cast<double*>(scratch);
  unsigned* tickets = reinterpret_cast<unsigned*>(static_cast<char*>(scratch) + size_t(kFinSegMax) * 2 * c * sizeof(double));
  if (S > 1) WFSP_CHECK_CUDA(cudaMemsetAsync(tickets, 0, 1024, st));
  bn_finalize_stats<<<dim3(unsigned(ceil_div(c, kCh)), unsigned(S)), dim3(kCh, kFinLanes), 0, st>>>(
      part, chunk_rows, n_rows, n_rows_dev, c, eps, momentum, running_mean, running_var, save_mean, save_invstd, inter,
      tickets);
  return WFSP_OK;
}

// rows per CTA pass of the apply kernels: 64 for large inputs; small (latency-bound) inputs are spread over
// ~4 CTAs per SM so that no thread walks more than a couple of dependent load batches
inline int apply_rows(int64_t rows) {
  int64_t r = ceil_div<int64_t>(rows > 0 ? rows : 1, int64_t(sm_count()) * 4);
  r = (r + 7) / 8 * 8;
  return int(r < 8 ? 8 : (r > kApplyRows ? kApplyRows : r));
}
inline unsigned apply_blocks(int64_t rows) {
  int64_t b = ceil_div<int64_t>(rows > 0 ? rows : 1, apply_rows(rows));
  const int64_t cap = int64_t(sm_count()) * 16;
  return unsigned(b > cap ? cap : b);
}

// float2 path of the apply kernels: even channel count and 8-byte aligned fp32 buffers (NULL = not used)
inline bool vec2_ok(int c, const void* a, const void* b) {
  return (c & 1) == 0 && (reinterpret_cast<uintptr_t>(a) & 7) == 0 && (reinterpret_cast<uintptr_t>(b) & 7) == 0;
}

// block shape of the apply kernels: x = channel pairs (a multiple of 32, at most 256), y = row lanes
inline dim3 apply_block(int c) {
  const int ppr = ((c + 7) & ~7) >> 1;
  int bx = (ppr + 31) / 32 * 32;
  if (bx > 256) bx = 256;
  int by = 256 / bx;
  return dim3(unsigned(bx), unsigned(by < 1 ? 1 : by), 1);
}

}  // namespace

// bn_stream.cu: the same passes fed by bulk copies (large row counts)
bool bn_stream_ok(int c, int arrays, const void* x, const void* dy);
int bn_stream_grid(int64_t n_rows, int chunk_rows);
int bn_stream_chunk_rows(int c, int arrays);
int bn_stream_fwd_apply(const float* x, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma, const float* beta,
                        const float* mean, const float* invstd, int relu, float* y, void* y16, const DropSpec& drop,
                        cudaStream_t st);
int bn_stream_bwd_partial(const float* x, const float* dy, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma,
                          const float* beta, const float* mean, const float* invstd, int relu, float* part, int* n_part,
                          const DropSpec& drop, cudaStream_t st);
int bn_stream_bwd_apply(const float* x, const float* dy, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma,
                        const float* beta, const float* mean, const float* invstd, const float* d_gamma, const float* d_beta,
                        int relu, float* dx, void* dx16, const DropSpec& drop, cudaStream_t st);
constexpr int64_t kStreamMinRows = 32768;  // (expected live) rows from which the bulk-copy fed passes are used
static int g_bn_stream = 1;                // wfsp_set_option "bn_stream": 0 = register-load kernels everywhere
void set_bn_stream(int v) { g_bn_stream = v; }
}  // namespace wfsp

using namespace wfsp;

extern "C" size_t wfsp_bn_workspace_bytes(int64_t n_rows, int c) {
  // sized for the finer of the two partial lists (backward: kRowsBwd rows per partial)
  return align_up(size_t(ceil_div<int64_t>(n_rows > 0 ? n_rows : 1, kRowsBwd)) * 2 * c * sizeof(float), 256) +
         fin_scratch_bytes(c);
}

extern "C" size_t wfsp_bn_partials_bytes(int64_t n_rows, int c) {
  return align_up(size_t(ceil_div<int64_t>(n_rows > 0 ? n_rows : 1, WFSP_BN_CHUNK_ROWS)) * 2 * c * sizeof(float), 256) +
         fin_scratch_bytes(c);
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
    WFSP_CHECK_CUDA(launch_cluster(bn_fwd_small, ceil_div((c + 7) & ~7, kCh), st, x, n_rows, n_rows_dev, c, gamma, beta,
                                   running_mean, running_var, momentum, eps, training, relu, y, y16, save_mean,
                                   save_invstd));
    count_launches(1);
    return WFSP_OK;
  }
  if (training) {
    if (workspace == nullptr || workspace_bytes < wfsp_bn_workspace_bytes(n_rows, c))
      return set_error(WFSP_EWORKSPACE, "batch-norm workspace too small");
    float* part = static_cast<float*>(workspace);
    dim3 grid(unsigned(ceil_div<int64_t>(n_rows, kRows)), unsigned(ceil_div(c, kCh)));
    bn_partial_stats<<<grid, dim3(32, kPartLanes), 0, st>>>(x, n_rows, n_rows_dev, c, part);
    char* scratch = static_cast<char*>(workspace) +
                    align_up(size_t(ceil_div<int64_t>(n_rows, kRows)) * 2 * c * sizeof(float), 256);
    if (int rc = launch_finalize(part, kRows, n_rows, n_rows_dev, c, eps, momentum, running_mean, running_var, save_mean,
                                 save_invstd, scratch, st))
      return rc;
    count_launches(2);
  } else {
    bn_eval_stats<<<ceil_div(c, 128), 128, 0, st>>>(c, eps, running_mean, running_var, save_mean, save_invstd);
    count_launches(1);
  }
  if (vec2_ok(c, x, y))
    bn_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, y, y16, PartStats{}, apply_rows(n_rows), DropSpec{});
  else
    bn_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, y, y16, PartStats{}, apply_rows(n_rows), DropSpec{});
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_fwd_stats_ex(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int64_t n_rows_hint, int c,
                                         const float* bn_partials, const float* gamma, const float* beta,
                                         float* running_mean, float* running_var, float momentum, float eps, int relu,
                                         float* y, void* y_bf16, float* save_mean, float* save_invstd,
                                         const wfsp_dropout* dropout, wfsp_stream_t stream) {
  const DropSpec drop = make_drop(dropout);
  WFSP_REQUIRE(n_rows >= 0 && c >= 1 && bn_partials != nullptr, "bad batch-norm arguments");
  WFSP_REQUIRE(y != nullptr || y_bf16 != nullptr, "batch norm needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* y16 = static_cast<__nv_bfloat16*>(y_bf16);
  const int64_t live = (n_rows_dev != nullptr && n_rows_hint > 0 && n_rows_hint < n_rows) ? n_rows_hint : n_rows;
  if (live <= kFoldRows && c <= 512) {
    // few chunks: every CTA of the apply kernel folds the partials of its channels itself -- ONE launch
    const PartStats ps{bn_partials, WFSP_BN_CHUNK_ROWS, eps, momentum, running_mean, running_var, save_mean, save_invstd};
    if (vec2_ok(c, x, y))
      bn_apply<true, true><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, nullptr, nullptr, relu, y, y16, ps, apply_rows(n_rows), drop);
    else
      bn_apply<false, true><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, nullptr, nullptr, relu, y, y16, ps, apply_rows(n_rows), drop);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  char* scratch = reinterpret_cast<char*>(const_cast<float*>(bn_partials)) +
                  align_up(size_t(ceil_div<int64_t>(n_rows, WFSP_BN_CHUNK_ROWS)) * 2 * c * sizeof(float), 256);
  if (int rc = launch_finalize(bn_partials, WFSP_BN_CHUNK_ROWS, n_rows, n_rows_dev, c, eps, momentum, running_mean,
                               running_var, save_mean, save_invstd, scratch, st))
    return rc;
  if (g_bn_stream && live >= kStreamMinRows && bn_stream_ok(c, 1, x, nullptr) && (y == nullptr || vec2_ok(c, y, y))) {
    count_launches(1);
    return bn_stream_fwd_apply(x, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, y, y16, drop, st);
  }
  if (vec2_ok(c, x, y))
    bn_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, y, y16, PartStats{}, apply_rows(n_rows), drop);
  else
    bn_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, y, y16, PartStats{}, apply_rows(n_rows), drop);
  count_launches(2);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_fwd_stats(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int64_t n_rows_hint, int c,
                                      const float* bn_partials, const float* gamma, const float* beta,
                                      float* running_mean, float* running_var, float momentum, float eps, int relu,
                                      float* y, void* y_bf16, float* save_mean, float* save_invstd,
                                      wfsp_stream_t stream) {
  return wfsp_bn_relu_fwd_stats_ex(x, n_rows, n_rows_dev, n_rows_hint, c, bn_partials, gamma, beta, running_mean, running_var,
                                   momentum, eps, relu, y, y_bf16, save_mean, save_invstd, nullptr, stream);
}

extern "C" int wfsp_bn_relu_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                int training, int relu, float* y, float* save_mean, float* save_invstd,
                                void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_bn_relu_fwd_x(x, n_rows, n_rows_dev, c, gamma, beta, running_mean, running_var, momentum, eps, training,
                            relu, y, nullptr, save_mean, save_invstd, workspace, workspace_bytes, stream);
}

extern "C" int wfsp_bn_relu_bwd_x_ex(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int64_t n_rows_hint,
                                     int c,
                                     const float* gamma, const float* beta, const float* save_mean,
                                     const float* save_invstd, int relu, float* dx, void* dx_bf16, float* d_gamma,
                                     float* d_beta, void* workspace, size_t workspace_bytes, const wfsp_dropout* dropout,
                                     wfsp_stream_t stream) {
  const DropSpec drop = make_drop(dropout);
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad batch-norm sizes");
  WFSP_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "batch norm backward needs at least one output");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* dx16 = static_cast<__nv_bfloat16*>(dx_bf16);
  if (n_rows == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_gamma, 0, size_t(c) * 4, st));
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_beta, 0, size_t(c) * 4, st));
    return WFSP_OK;
  }
  const int64_t live = (n_rows_dev != nullptr && n_rows_hint > 0 && n_rows_hint < n_rows) ? n_rows_hint : n_rows;
  if (live <= kClusterRows && n_rows <= kSmallRows * 4) {
    WFSP_CHECK_CUDA(launch_cluster(bn_bwd_small, ceil_div((c + 7) & ~7, kCh), st, x, dy, n_rows, n_rows_dev, c, gamma, beta,
                                   save_mean, save_invstd, relu, dx, dx16, d_gamma, d_beta, drop));
    count_launches(1);
    return WFSP_OK;
  }
  if (workspace == nullptr || workspace_bytes < wfsp_bn_workspace_bytes(n_rows, c))
    return set_error(WFSP_EWORKSPACE, "batch-norm workspace too small");
  float* part = static_cast<float*>(workspace);
  if (g_bn_stream && live >= kStreamMinRows && n_rows >= int64_t(sm_count()) * kRowsBwd && bn_stream_ok(c, 2, x, dy) &&
      (dx == nullptr || vec2_ok(c, dx, dx))) {
    int n_part = 0;
    if (int rc = bn_stream_bwd_partial(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, part, &n_part, drop, st))
      return rc;
    bn_bwd_finalize<<<ceil_div(c, kCh), dim3(kCh, kFinLanes), 0, st>>>(part, 1, n_rows, n_rows_dev, c, d_gamma, d_beta, nullptr,
                                                                      nullptr, n_part);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return bn_stream_bwd_apply(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, relu, dx, dx16, drop, st);
  }
  const unsigned pgrid = unsigned(ceil_div<int64_t>(n_rows, kRowsBwd));
  if (vec2_ok(c, x, dy))
    bn_bwd_partial<true><<<pgrid, apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, part, drop);
  else
    bn_bwd_partial<false><<<pgrid, apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, relu, part, drop);
  bn_bwd_finalize<<<ceil_div(c, kCh), dim3(kCh, kFinLanes), 0, st>>>(part, kRowsBwd, n_rows, n_rows_dev, c, d_gamma, d_beta, nullptr, nullptr, -1);
  if (vec2_ok(c, x, dy) && vec2_ok(c, dx, dx))
    bn_bwd_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, relu, dx, dx16, apply_rows(n_rows), nullptr, 0, nullptr, nullptr, drop);
  else
    bn_bwd_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, relu, dx, dx16, apply_rows(n_rows), nullptr, 0, nullptr, nullptr, drop);
  count_launches(3);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_bwd_x(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int64_t n_rows_hint,
                                  int c, const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_invstd, int relu, float* dx, void* dx_bf16, float* d_gamma,
                                  float* d_beta, void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_bn_relu_bwd_x_ex(x, dy, n_rows, n_rows_dev, n_rows_hint, c, gamma, beta, save_mean, save_invstd, relu, dx, dx_bf16,
                               d_gamma, d_beta, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int wfsp_bn_relu_bwd_parts(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev,
                                      int64_t n_rows_hint, int c, const float* gamma, const float* beta,
                                      const float* save_mean, const float* save_invstd, int relu,
                                      const float* bwd_partials, float* dx, void* dx_bf16, float* d_gamma, float* d_beta,
                                      wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1 && bwd_partials != nullptr && d_gamma != nullptr && d_beta != nullptr,
               "bad batch-norm arguments");
  WFSP_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "batch norm backward needs at least one output");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* dx16 = static_cast<__nv_bfloat16*>(dx_bf16);
  if (n_rows == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_gamma, 0, size_t(c) * 4, st));
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_beta, 0, size_t(c) * 4, st));
    return WFSP_OK;
  }
  const int64_t live = (n_rows_dev != nullptr && n_rows_hint > 0 && n_rows_hint < n_rows) ? n_rows_hint : n_rows;
  const bool v2 = vec2_ok(c, x, dy) && vec2_ok(c, dx, dx);
  if (live <= kFoldRows && c <= 512) {  // ONE launch: every CTA folds the few partials of its channels itself
    if (v2)
      bn_bwd_apply<true, true><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, nullptr, nullptr, relu, dx, dx16, apply_rows(n_rows), bwd_partials, WFSP_BN_CHUNK_ROWS, d_gamma, d_beta, DropSpec{});
    else
      bn_bwd_apply<false, true><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, nullptr, nullptr, relu, dx, dx16, apply_rows(n_rows), bwd_partials, WFSP_BN_CHUNK_ROWS, d_gamma, d_beta, DropSpec{});
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  const int64_t chunks = ceil_div<int64_t>(n_rows, WFSP_BN_CHUNK_ROWS);
  char* scratch = reinterpret_cast<char*>(const_cast<float*>(bwd_partials)) + align_up(size_t(chunks) * 2 * c * sizeof(float), 256);
  const int S = fin_segments(chunks);
  double* inter = reinterpret_cast<double*>(scratch);
  unsigned* tickets = reinterpret_cast<unsigned*>(scratch + size_t(kFinSegMax) * 2 * c * sizeof(double));
  if (S > 1) WFSP_CHECK_CUDA(cudaMemsetAsync(tickets, 0, 1024, st));
  bn_bwd_finalize<<<dim3(unsigned(ceil_div(c, kCh)), unsigned(S)), dim3(kCh, kFinLanes), 0, st>>>(
      bwd_partials, WFSP_BN_CHUNK_ROWS, n_rows, n_rows_dev, c, d_gamma, d_beta, inter, tickets, -1);
  if (v2)
    bn_bwd_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, relu, dx, dx16, apply_rows(n_rows), nullptr, 0, nullptr, nullptr, DropSpec{});
  else
    bn_bwd_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, st>>>(x, dy, n_rows, n_rows_dev, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, relu, dx, dx16, apply_rows(n_rows), nullptr, 0, nullptr, nullptr, DropSpec{});
  count_launches(2);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_bn_relu_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c,
                                const float* gamma, const float* beta, const float* save_mean,
                                const float* save_invstd, int relu, float* dx, float* d_gamma, float* d_beta,
                                void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_bn_relu_bwd_x(x, dy, n_rows, n_rows_dev, 0, c, gamma, beta, save_mean, save_invstd, relu, dx, nullptr, d_gamma,
                            d_beta, workspace, workspace_bytes, stream);
}

namespace wfsp {
namespace {
// column sums of the LIVE rows of x [cap, c] (bias gradient of a convolution on the graph path): per 256-row chunk
// partials (row lanes combined in lane order), folded in chunk order by bn_bwd_finalize -- deterministic
__global__ void __launch_bounds__(256) col_sum_partial(const float* __restrict__ x, int64_t n_cap, const int32_t* __restrict__ n_dev,
                                                       int c, float* __restrict__ part /* [chunks][2][c], slot 1 = 0 */) {
  __shared__ float red[8][32];
  const int64_t n = live_rows(n_cap, n_dev);
  const int64_t r0 = int64_t(blockIdx.x) * kRowsBwd;
  if (r0 >= n) return;
  const int64_t r_end = r0 + kRowsBwd < n ? r0 + kRowsBwd : n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int ch0 = 0; ch0 < c; ch0 += 32) {
    const int ch = ch0 + tx;
    float s = 0.f;
    if (ch < c)
      for (int64_t r = r0 + ty; r < r_end; r += 8) s += x[r * c + ch];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && ch < c) {
      float t = 0.f;
      for (int l = 0; l < 8; ++l) t += red[l][tx];
      part[(int64_t(blockIdx.x) * 2 + 0) * c + ch] = t;
      part[(int64_t(blockIdx.x) * 2 + 1) * c + ch] = 0.f;
    }
    __syncthreads();
  }
}
}  // namespace
}  // namespace wfsp

extern "C" int wfsp_col_sum(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, float* out, void* workspace,
                            size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1 && out != nullptr, "bad column-sum arguments");
  cudaStream_t st = as_stream(stream);
  if (n_rows == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(out, 0, size_t(c) * 4, st));
    return WFSP_OK;
  }
  if (workspace == nullptr || workspace_bytes < wfsp_bn_workspace_bytes(n_rows, c))
    return set_error(WFSP_EWORKSPACE, "column-sum workspace too small");
  float* part = static_cast<float*>(workspace);
  // the fold kernel of BatchNorm backward adds slot 0 into its "d_beta" output (= the column sums) and slot 1 (zeros)
  // into its "d_gamma" output, which goes to the scratch area behind the partial list
  float* unused = reinterpret_cast<float*>(static_cast<char*>(workspace) +
                                           align_up(size_t(ceil_div<int64_t>(n_rows, kRowsBwd)) * 2 * c * sizeof(float), 256));
  col_sum_partial<<<unsigned(ceil_div<int64_t>(n_rows, kRowsBwd)), 256, 0, st>>>(x, n_rows, n_rows_dev, c, part);
  bn_bwd_finalize<<<ceil_div(c, kCh), dim3(kCh, kFinLanes), 0, st>>>(part, kRowsBwd, n_rows, n_rows_dev, c, unused, out, nullptr,
                                                                    nullptr, -1);
  count_launches(2);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

namespace wfsp {
namespace {
__global__ void __launch_bounds__(256) dropout_factors_kernel(const DropSpec drop, int64_t total, float* __restrict__ out) {
  const unsigned long long key = drop_key(drop);
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < total; i += int64_t(gridDim.x) * 256)
    out[i] = drop.p > 0.f ? drop_factor(key, drop.p, drop.scale, (unsigned long long)i) : 1.f;
}
}  // namespace
}  // namespace wfsp

extern "C" int wfsp_dropout_factors(const wfsp_dropout* dropout, int64_t n_rows, int c, float* factors, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1 && factors != nullptr, "bad dropout arguments");
  WFSP_REQUIRE(dropout == nullptr || (dropout->p >= 0.f && dropout->p < 1.f), "dropout probability must be in [0, 1)");
  const int64_t total = n_rows * c;
  if (total == 0) return WFSP_OK;
  int64_t blocks = ceil_div<int64_t>(total, 256);
  if (blocks > int64_t(sm_count()) * 8) blocks = int64_t(sm_count()) * 8;
  dropout_factors_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(make_drop(dropout), total, factors);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

// Activation-only blocks (a convolution followed by ReLU / nothing, no BatchNorm): the same streaming
// kernels with the normalisation switched off.
extern "C" int wfsp_act_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, int relu, float* y,
                            void* y_bf16, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad sizes");
  WFSP_REQUIRE(y != nullptr || y_bf16 != nullptr, "needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  if (vec2_ok(c, x, y))
    bn_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, as_stream(stream)>>>(
        x, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, relu, y, static_cast<__nv_bfloat16*>(y_bf16), PartStats{}, apply_rows(n_rows), DropSpec{});
  else
    bn_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, as_stream(stream)>>>(
        x, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, relu, y, static_cast<__nv_bfloat16*>(y_bf16), PartStats{}, apply_rows(n_rows), DropSpec{});
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_act_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c, int relu,
                            float* dx, void* dx_bf16, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1, "bad sizes");
  WFSP_REQUIRE(dx != nullptr || dx_bf16 != nullptr, "needs at least one output");
  if (n_rows == 0) return WFSP_OK;
  if (vec2_ok(c, x, dy) && vec2_ok(c, dx, dx))
    bn_bwd_apply<true, false><<<apply_blocks(n_rows), apply_block(c), 0, as_stream(stream)>>>(
        x, dy, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, relu, dx,
        static_cast<__nv_bfloat16*>(dx_bf16), apply_rows(n_rows), nullptr, 0, nullptr, nullptr, DropSpec{});
  else
    bn_bwd_apply<false, false><<<apply_blocks(n_rows), apply_block(c), 0, as_stream(stream)>>>(
        x, dy, n_rows, n_rows_dev, c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, relu, dx,
        static_cast<__nv_bfloat16*>(dx_bf16), apply_rows(n_rows), nullptr, 0, nullptr, nullptr, DropSpec{});
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
