// Dense classification head + loss of the PSD classifier in three launches.
//
// The reference's SPConvNet flattens the ToDense output and applies LinearBlock = Linear(4480,116) .
// Linear(116,3) with no activation in between (src/models/ConvBlocks.py:82-102, src/models/SPConvNet.py:67-68),
// and LitPSD.training_step takes CrossEntropyLoss (mean) of the result (src/engineering/LitPSD.py:94-104).
// At the reference's batch of 64 events that is ~25 library kernels (split-K SGEMM + reduce, bias, softmax,
// nll, their backward, bias-gradient reductions) of 2-40 us each -- more than 40 % of a whole training step
// once the sparse layers run in a few microseconds each.  Here:
//
//   head_l1_partial   h1 partial sums over K-chunks of the 4480-wide reduction (all SMs busy), fp32 FMA
//   head_tail         one CTA: reduce the partials (+bias) -> h1; logits; softmax / cross-entropy (mean);
//                     and the whole small half of the backward pass: dlogits, dW2, db2, dh1
//   head_bwd          dW1 = dh1^T x, dx = dh1 W1, db1 per K-chunk (scaled by the incoming loss gradient)
//
// Same arithmetic as the torch modules (fp32 products and sums; summation order differs).
#include "common.cuh"

namespace wfsp {
namespace {

constexpr int kKC = 128;   // reduction chunk of the first layer per CTA
constexpr int kHT = 32;    // h1 columns per CTA
constexpr int kBT = 32;    // batch rows per CTA

// grid (splits, ceil(H1/32), ceil(B/32)), block 256.  partial[split][b][h]
__global__ void __launch_bounds__(256) head_l1_partial(const float* __restrict__ x, const float* __restrict__ w1, int B, int K0,
                                                       int H1, float* __restrict__ partial) {
  __shared__ float xs[kBT][kKC + 1];
  __shared__ float ws[kHT][kKC + 1];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * kKC, h0 = blockIdx.y * kHT, b0 = blockIdx.z * kBT;
  // thread (row = tid / 32 + 8 j, 4 consecutive k per step): coalesced 128-byte row segments, no div / mod
  const int fr = tid >> 5, fk = (tid & 31) * 4;
#pragma unroll
  for (int j = 0; j < kBT / 8; ++j) {
    const int r = fr + 8 * j;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = fk + e;
      xs[r][k] = (b0 + r < B && k0 + k < K0) ? x[int64_t(b0 + r) * K0 + k0 + k] : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < kHT / 8; ++j) {
    const int r = fr + 8 * j;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = fk + e;
      ws[r][k] = (h0 + r < H1 && k0 + k < K0) ? w1[int64_t(h0 + r) * K0 + k0 + k] : 0.f;
    }
  }
  __syncthreads();
  const int th = tid & 31, tb = tid >> 5;  // column h0 + th, rows b0 + tb*4 .. +3
  float acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kKC; ++k) {
    const float wv = ws[th][k];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += xs[tb * 4 + i][k] * wv;
  }
  if (h0 + th < H1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + tb * 4 + i;
      if (b < B) partial[(int64_t(blockIdx.x) * B + b) * H1 + h0 + th] = acc[i];
    }
  }
}

// h1 = sum of the K-chunk partials + b1.  Four lanes per element take every fourth chunk and are combined by
// shuffles in a fixed order (deterministic); consecutive lane groups = consecutive elements (coalesced).
__global__ void __launch_bounds__(256) head_l1_reduce(const float* __restrict__ partial, int splits, const float* __restrict__ b1,
                                                      int B, int H1, float* __restrict__ h1, unsigned* __restrict__ tail_ticket) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t == 0 && tail_ticket != nullptr) *tail_ticket = 0u;  // the multi-CTA tail launched behind this kernel counts from zero
  const int i = t >> 2, q = t & 3;
  float s = 0.f;
  if (i < B * H1)
#pragma unroll 4
    for (int sp = q; sp < splits; sp += 4) s += partial[int64_t(sp) * B * H1 + i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (q == 0 && i < B * H1) h1[i] = s + (b1 ? b1[i % H1] : 0.f);
}

// one CTA of 1024 threads: everything that is O(batch x hidden) or smaller.  h1, w2 and dlogits are staged in
// shared memory once (coalesced), so the serial inner loops run on chip instead of on L2 latency.
// dynamic smem: h1 [B][H1+1] | w2 [C][H1+1] | dlogits [B][C] | logits [B][C]
__global__ void __launch_bounds__(1024) head_tail(const float* __restrict__ w2, const float* __restrict__ b2,
                                                  const int64_t* __restrict__ labels, int B, int H1, int C,
                                                  const float* __restrict__ h1, float* __restrict__ logits, float* __restrict__ loss,
                                                  float* __restrict__ dlogits, float* __restrict__ dh1, float* __restrict__ dw2,
                                                  float* __restrict__ db2) {
  extern __shared__ float sm[];
  __shared__ float s_red[32];
  const int tid = threadIdx.x, nt = blockDim.x, hp = H1 + 1;
  float* s_h1 = sm;
  float* s_w2 = s_h1 + B * hp;
  float* s_dl = s_w2 + C * hp;
  float* s_lg = s_dl + B * C;
#pragma unroll 1
  for (int i = tid; i < B * H1; i += nt) s_h1[(i / H1) * hp + i % H1] = h1[i];
#pragma unroll 1
  for (int i = tid; i < C * H1; i += nt) s_w2[(i / H1) * hp + i % H1] = w2[i];
  __syncthreads();
  // logits = h1 w2^T + b2
#pragma unroll 1
  for (int i = tid; i < B * C; i += nt) {
    const int b = i / C, c = i % C;
    float s = b2 ? b2[c] : 0.f;
    const float* hr = s_h1 + b * hp;
    const float* wr = s_w2 + c * hp;
#pragma unroll 4
    for (int h = 0; h < H1; ++h) s += hr[h] * wr[h];
    s_lg[i] = s;
    logits[i] = s;
  }
  __syncthreads();
  // cross entropy (mean over the batch) and dlogits = (softmax - onehot) / B
  float lsum = 0.f;
#pragma unroll 1
  for (int b = tid; b < B; b += nt) {
    const float* lr = s_lg + b * C;
    float m = lr[0];
#pragma unroll 1
    for (int c = 1; c < C; ++c) m = fmaxf(m, lr[c]);
    float z = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) z += expf(lr[c] - m);
    const float lse = m + logf(z);
    const int y = int(labels[b]);
    lsum += lse - lr[y];
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
      const float d = (expf(lr[c] - lse) - (c == y ? 1.f : 0.f)) / float(B);
      s_dl[b * C + c] = d;
      dlogits[int64_t(b) * C + c] = d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll 1
    for (int w = 0; w < (nt + 31) / 32; ++w) t += s_red[w];
    *loss = t / float(B);
  }
  // dW2[c][h] = sum_b dlogits[b][c] h1[b][h];  db2[c] = sum_b dlogits[b][c]
#pragma unroll 1
  for (int i = tid; i < C * H1; i += nt) {
    const int c = i / H1, h = i % H1;
    float s = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) s += s_dl[b * C + c] * s_h1[b * hp + h];
    dw2[i] = s;
  }
#pragma unroll 1
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
#pragma unroll 1
    for (int b = 0; b < B; ++b) s += s_dl[b * C + c];
    db2[c] = s;
  }
  // dh1[b][h] = sum_c dlogits[b][c] w2[c][h]
#pragma unroll 1
  for (int i = tid; i < B * H1; i += nt) {
    const int b = i / H1, h = i % H1;
    float s = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) s += s_dl[b * C + c] * s_w2[c * hp + h];
    dh1[i] = s;
  }
}

// The same tail for ANY batch: CTA t owns batch rows [16 t, 16 t + 16).  Per-row results (logits, dlogits, dh1) are
// final per CTA; the batch reductions (loss, dW2, db2) go through per-tile partials that the LAST CTA to finish (ticket)
// adds in tile order -- deterministic, one launch.  h1 = x w1^T + b1 comes from the caller (a plain GEMM: library).
// dynamic smem: h1 [TB][H1+1] | w2 [C][H1+1] | dlogits [TB][C] | logits [TB][C];  part: [tiles][C*H1 + C + 1]
constexpr int kTailRows = 16;
__global__ void __launch_bounds__(256) head_tail_tiles(const float* __restrict__ w2, const float* __restrict__ b2,
                                                       const int64_t* __restrict__ labels, int B, int H1, int C,
                                                       const float* __restrict__ h1, float* __restrict__ logits,
                                                       float* __restrict__ loss, float* __restrict__ dlogits,
                                                       float* __restrict__ dh1, float* __restrict__ dw2, float* __restrict__ db2,
                                                       float* __restrict__ part, unsigned* __restrict__ ticket) {
  extern __shared__ float sm[];
  __shared__ float s_red[16];
  __shared__ unsigned s_ticket;
  const int tid = threadIdx.x, nt = blockDim.x, hp = H1 + 1;
  const int b0 = blockIdx.x * kTailRows, nb = min(kTailRows, B - b0), tiles = gridDim.x;
  float* s_h1 = sm;
  float* s_w2 = s_h1 + kTailRows * hp;
  float* s_dl = s_w2 + C * hp;
  float* s_lg = s_dl + kTailRows * C;
  for (int i = tid; i < nb * H1; i += nt) s_h1[(i / H1) * hp + i % H1] = h1[int64_t(b0) * H1 + i];
  for (int i = tid; i < C * H1; i += nt) s_w2[(i / H1) * hp + i % H1] = w2[i];
  __syncthreads();
  for (int i = tid; i < nb * C; i += nt) {
    const int b = i / C, c = i % C;
    float s = b2 ? b2[c] : 0.f;
    const float* hr = s_h1 + b * hp;
    const float* wr = s_w2 + c * hp;
#pragma unroll 4
    for (int h = 0; h < H1; ++h) s += hr[h] * wr[h];
    s_lg[i] = s;
    logits[int64_t(b0) * C + i] = s;
  }
  __syncthreads();
  float lsum = 0.f;
  for (int b = tid; b < nb; b += nt) {
    const float* lr = s_lg + b * C;
    float m = lr[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, lr[c]);
    float z = 0.f;
    for (int c = 0; c < C; ++c) z += expf(lr[c] - m);
    const float lse = m + logf(z);
    const int y = int(labels[b0 + b]);
    lsum += lse - lr[y];
    for (int c = 0; c < C; ++c) {
      const float d = (expf(lr[c] - lse) - (c == y ? 1.f : 0.f)) / float(B);
      s_dl[b * C + c] = d;
      dlogits[int64_t(b0 + b) * C + c] = d;
    }
  }
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = lsum;
  __syncthreads();
  float* my = part + int64_t(blockIdx.x) * (C * H1 + C + 1);
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < nt / 32; ++w) t += s_red[w];
    my[C * H1 + C] = t;
  }
  for (int i = tid; i < C * H1; i += nt) {
    const int c = i / H1, h = i % H1;
    float s = 0.f;
#pragma unroll 4
    for (int b = 0; b < nb; ++b) s += s_dl[b * C + c] * s_h1[b * hp + h];
    my[i] = s;
  }
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
    for (int b = 0; b < nb; ++b) s += s_dl[b * C + c];
    my[C * H1 + c] = s;
  }
  for (int i = tid; i < nb * H1; i += nt) {
    const int b = i / H1, h = i % H1;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += s_dl[b * C + c] * s_w2[c * hp + h];
    dh1[int64_t(b0) * H1 + i] = s;
  }
  // last CTA: the batch reductions, in tile order
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != unsigned(tiles - 1)) return;
  __threadfence();
  if (tid == 0) *ticket = 0u;
  const int per = C * H1 + C + 1;
  for (int i = tid; i < per; i += nt) {
    float s = 0.f;
    for (int t = 0; t < tiles; ++t) s += __ldcg(part + int64_t(t) * per + i);
    if (i < C * H1) dw2[i] = s;
    else if (i < C * H1 + C) db2[i - C * H1] = s;
    else *loss = s / float(B);
  }
}

constexpr int kKB = 32;    // K-chunk of the backward kernel
constexpr int kBB = 32;    // batch rows per tile of the backward kernel
constexpr int kMaxH1 = 128;

// grid (ceil(K0/32) + 1), block 256.  Chunk CTAs: dW1[:, chunk] and dx[:, chunk]; the extra last CTA: db1 and the
// scaling of the small gradients (dW2, db2) by the incoming loss gradient.
__global__ void __launch_bounds__(256) head_bwd(const float* __restrict__ x, const float* __restrict__ w1,
                                                const float* __restrict__ dh1, const float* __restrict__ dw2_in,
                                                const float* __restrict__ db2_in, const float* __restrict__ go, int B, int K0,
                                                int H1, int C, float* __restrict__ dx, float* __restrict__ dw1,
                                                float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
  const int tid = threadIdx.x;
  const float g = go ? *go : 1.f;
  const int chunks = (K0 + kKB - 1) / kKB;
  if (int(blockIdx.x) == chunks) {
    for (int h = tid; h < H1; h += 256) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dh1[int64_t(b) * H1 + h];
      if (db1) db1[h] = s * g;
    }
    for (int i = tid; i < C * H1; i += 256) dw2[i] = dw2_in[i] * g;
    for (int i = tid; i < C; i += 256) db2[i] = db2_in[i] * g;
    return;
  }
  __shared__ float dhs[kBB][kMaxH1 + 1];  // dh1 tile [32 rows][H1]
  __shared__ float xs[kBB][kKB + 1];      // x tile   [32 rows][32 k]
  __shared__ float w1s[kMaxH1][kKB + 1];  // W1 chunk [H1][32 k]
  const int k0 = blockIdx.x * kKB;
  for (int h = tid >> 5; h < H1; h += 8) {  // a warp reads one 128-byte row segment of W1
    const int k = tid & 31;
    w1s[h][k] = (k0 + k < K0) ? w1[int64_t(h) * K0 + k0 + k] : 0.f;
  }
  // dW1 accumulators: thread (tk = tid % 32, th = tid / 32) owns column k0 + tk, rows h = th + 8 j
  const int tk = tid & 31, th = tid >> 5;
  float accw[kMaxH1 / 8];
#pragma unroll
  for (int j = 0; j < kMaxH1 / 8; ++j) accw[j] = 0.f;
  for (int b0 = 0; b0 < B; b0 += kBB) {
    __syncthreads();
    for (int r = tid >> 5; r < kBB; r += 8) {
      for (int h = tid & 31; h < kMaxH1; h += 32)
        dhs[r][h] = (b0 + r < B && h < H1) ? dh1[int64_t(b0 + r) * H1 + h] : 0.f;
      const int k = tid & 31;
      xs[r][k] = (b0 + r < B && k0 + k < K0) ? x[int64_t(b0 + r) * K0 + k0 + k] : 0.f;
    }
    __syncthreads();
    // dW1[h][k] += sum_r dhs[r][h] xs[r][k]
    for (int r = 0; r < kBB; ++r) {
      const float xv = xs[r][tk];
#pragma unroll
      for (int j = 0; j < kMaxH1 / 8; ++j) accw[j] += dhs[r][th + 8 * j] * xv;   // columns past H1 hold zeros / stale: masked on store
    }
    // dx[b][k] = sum_h dhs[b][h] w1s[h][k]: thread owns column tk, rows th + 8 j (8 rows)
    if (dx) {
#pragma unroll
      for (int j = 0; j < kBB / 8; ++j) {
        const int r = th + 8 * j;
        float s = 0.f;
        for (int h = 0; h < H1; ++h) s += dhs[r][h] * w1s[h][tk];
        if (b0 + r < B && k0 + tk < K0) dx[int64_t(b0 + r) * K0 + k0 + tk] = s * g;
      }
    }
  }
  if (k0 + tk < K0) {
#pragma unroll
    for (int j = 0; j < kMaxH1 / 8; ++j) {
      const int h = th + 8 * j;
      if (h < H1) dw1[int64_t(h) * K0 + k0 + tk] = accw[j] * g;
    }
  }
}

// the small half of the backward pass: dh1_scaled = dh1 * g (operand of the two library GEMMs dW1 = dh1^T x,
// dx = dh1 W1), db1 = column sums * g, dW2 / db2 = forward's values * g.  grid = ceil(H1 / 8) CTAs of 256.
__global__ void __launch_bounds__(256) head_bwd_small(const float* __restrict__ dh1, const float* __restrict__ dw2_in,
                                                      const float* __restrict__ db2_in, const float* __restrict__ go, int B,
                                                      int H1, int C, float* __restrict__ dh1_scaled, float* __restrict__ db1,
                                                      float* __restrict__ dw2, float* __restrict__ db2) {
  const float g = go ? *go : 1.f;
  const int tid = threadIdx.x, gtid = blockIdx.x * 256 + tid, gn = gridDim.x * 256;
  for (int i = gtid; i < B * H1; i += gn) dh1_scaled[i] = dh1[i] * g;
  for (int i = gtid; i < C * H1; i += gn) dw2[i] = dw2_in[i] * g;
  for (int i = gtid; i < C; i += gn) db2[i] = db2_in[i] * g;
  if (db1) {  // warp w of this CTA sums column blockIdx.x * 8 + w over the batch
    const int h = blockIdx.x * 8 + (tid >> 5);
    if (h < H1) {
      float s = 0.f;
      for (int b = tid & 31; b < B; b += 32) s += dh1[int64_t(b) * H1 + h];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) db1[h] = s * g;
    }
  }
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" size_t wfsp_head_tail_workspace_bytes(int batch, int h1, int n_class);
extern "C" size_t wfsp_head_workspace_bytes(int batch, int k0, int h1) {
  const size_t splits = size_t((k0 + kKC - 1) / kKC);
  // split-K partials of Linear-1 | per-tile partials of the tail (for up to 64 classes) | the tail's ticket
  return align_up(splits * size_t(batch) * h1 * sizeof(float), 256) + wfsp_head_tail_workspace_bytes(batch, h1, 64) + 256;
}

extern "C" int wfsp_head_ce_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                const int64_t* labels, int batch, int k0, int h1_dim, int n_class, float* h1, float* logits,
                                float* loss, float* dlogits, float* dh1, float* dw2, float* db2, void* workspace,
                                size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 1 && k0 >= 1 && h1_dim >= 1 && h1_dim <= kMaxH1 && n_class >= 1 && n_class <= 64,
               "head sizes outside the supported range (hidden <= 128, classes <= 64)");
  WFSP_REQUIRE(x && w1 && w2 && labels && h1 && logits && loss && dlogits && dh1 && dw2 && db2, "null argument");
  if (workspace == nullptr || workspace_bytes < wfsp_head_workspace_bytes(batch, k0, h1_dim))
    return set_error(WFSP_EWORKSPACE, "head workspace too small");
  cudaStream_t st = as_stream(stream);
  const int splits = (k0 + kKC - 1) / kKC;
  float* partial = static_cast<float*>(workspace);
  dim3 grid(unsigned(splits), unsigned((h1_dim + kHT - 1) / kHT), unsigned((batch + kBT - 1) / kBT));
  head_l1_partial<<<grid, 256, 0, st>>>(x, w1, batch, k0, h1_dim, partial);
  // the tail runs as ceil(batch / 16) CTAs (logits / loss terms / dlogits / dh1 are per-row work; the batch reductions
  // are folded by the last CTA in tile order): the one-CTA version was 12 us of four dependent phases at 64 events
  char* w8 = static_cast<char*>(workspace);
  const size_t off_tail = align_up(size_t(splits) * batch * h1_dim * sizeof(float), 256);
  float* tail_part = reinterpret_cast<float*>(w8 + off_tail);
  unsigned* ticket = reinterpret_cast<unsigned*>(w8 + off_tail + wfsp_head_tail_workspace_bytes(batch, h1_dim, 64));
  head_l1_reduce<<<unsigned((batch * h1_dim * 4 + 255) / 256), 256, 0, st>>>(partial, splits, b1, batch, h1_dim, h1, ticket);
  const size_t tail_smem = (size_t(kTailRows) * (h1_dim + 1) + size_t(n_class) * (h1_dim + 1) + size_t(2) * kTailRows * n_class) * sizeof(float);
  WFSP_CHECK_CUDA(cudaFuncSetAttribute(head_tail_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  head_tail_tiles<<<unsigned((batch + kTailRows - 1) / kTailRows), 256, tail_smem, st>>>(
      w2, b2, labels, batch, h1_dim, n_class, h1, logits, loss, dlogits, dh1, dw2, db2, tail_part, ticket);
  count_launches(3);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" size_t wfsp_head_tail_workspace_bytes(int batch, int h1, int n_class) {
  const size_t tiles = size_t((batch + kTailRows - 1) / kTailRows);
  return align_up(tiles * (size_t(n_class) * h1 + n_class + 1) * sizeof(float), 256);
}

extern "C" int wfsp_head_ce_tail(const float* h1, const float* w2, const float* b2, const int64_t* labels, int batch,
                                 int h1_dim, int n_class, float* logits, float* loss, float* dlogits, float* dh1, float* dw2,
                                 float* db2, void* workspace, size_t workspace_bytes, unsigned* ticket, wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 1 && h1_dim >= 1 && h1_dim <= kMaxH1 && n_class >= 1 && n_class <= 64,
               "head sizes outside the supported range (hidden <= 128, classes <= 64)");
  WFSP_REQUIRE(h1 && w2 && labels && logits && loss && dlogits && dh1 && dw2 && db2 && ticket, "null argument");
  if (workspace == nullptr || workspace_bytes < wfsp_head_tail_workspace_bytes(batch, h1_dim, n_class))
    return set_error(WFSP_EWORKSPACE, "head tail workspace too small");
  const size_t smem = (size_t(kTailRows) * (h1_dim + 1) + size_t(n_class) * (h1_dim + 1) + size_t(2) * kTailRows * n_class) * sizeof(float);
  WFSP_CHECK_CUDA(cudaFuncSetAttribute(head_tail_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  head_tail_tiles<<<unsigned((batch + kTailRows - 1) / kTailRows), 256, smem, as_stream(stream)>>>(
      w2, b2, labels, batch, h1_dim, n_class, h1, logits, loss, dlogits, dh1, dw2, db2, static_cast<float*>(workspace), ticket);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_head_bwd(const float* x, const float* w1, const float* dh1, const float* dw2_in, const float* db2_in,
                             const float* grad_out, int batch, int k0, int h1_dim, int n_class, float* dx, float* dw1,
                             float* db1, float* dw2, float* db2, wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 1 && k0 >= 1 && h1_dim >= 1 && h1_dim <= kMaxH1 && n_class >= 1, "head sizes outside the supported range");
  WFSP_REQUIRE(x && w1 && dh1 && dw2_in && db2_in && dw1 && dw2 && db2, "null argument");
  const int chunks = (k0 + kKB - 1) / kKB;
  head_bwd<<<unsigned(chunks + 1), 256, 0, as_stream(stream)>>>(x, w1, dh1, dw2_in, db2_in, grad_out, batch, k0, h1_dim, n_class,
                                                               dx, dw1, db1, dw2, db2);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_head_bwd_small(const float* dh1, const float* dw2_in, const float* db2_in, const float* grad_out, int batch,
                                   int h1_dim, int n_class, float* dh1_scaled, float* db1, float* dw2, float* db2,
                                   wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 1 && h1_dim >= 1 && n_class >= 1, "bad head sizes");
  WFSP_REQUIRE(dh1 && dw2_in && db2_in && dh1_scaled && dw2 && db2, "null argument");
  head_bwd_small<<<unsigned((h1_dim + 7) / 8), 256, 0, as_stream(stream)>>>(dh1, dw2_in, db2_in, grad_out, batch, h1_dim, n_class,
                                                                            dh1_scaled, db1, dw2, db2);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
