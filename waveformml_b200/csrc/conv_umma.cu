// tcgen05 (5th-gen tensor core) gather-GEMM kernels, WFSP_MATH_BF16: bf16 operands, fp32
// accumulation in tensor memory.  They replace upstream indiceConv / indiceConvBackward
// (reference call sites src/models/SPConvBlocks.py:498-502 etc.).
//
// Both kernels are warp-specialised (544 threads):
//   warps 0-15 producers: issue 16-byte cp.async copies of gathered bf16 rows straight into
//              128-byte-swizzled shared-memory tiles, then cp.async.mbarrier.arrive.noinc on the stage's
//              "full" barrier and move on -- they never wait for data, so up to `stages` slices are in
//              flight per CTA.  Sixteen warps (four per scheduler) because a lone producer warp per
//              scheduler was instruction-issue bound (ncu: ~0.27 IPC);
//   warp 16    MMA issuer.  The whole warp walks the pipeline with converged control flow and ONE ELECTED lane
//              issues the four tcgen05.mma (K=16 each, accumulating in TMEM) of a slice and tcgen05.commit's to
//              the stage's "free" barrier.  (A lone `lane == 0` branch made the compiler wrap every tcgen05
//              instruction in a loop over the active lanes and move each descriptor through R2UR: ~100 cycles
//              per MMA, which WAS the cost of a small launch -- cycle trace in DESIGN.md.)
//   warp 17    apply kernel only: one cp.async.bulk per stage brings the pre-swizzled weight slice (expect_tx);
//   warps 0-7  epilogue after the last commit: tcgen05.ld -> registers -> shared-memory tile -> coalesced
//              global rows (warps w and w+4 share TMEM lane quarter w and take alternate column chunks).
//
// apply kernel (forward, dgrad, inverse forward, inverse dgrad) -- output-stationary implicit GEMM:
//   a CTA owns RB x 128 destination rows and walks (active kernel offset k) x (64-channel slice):
//   A tile = rows nbr[r][k] of the bf16 activation copy (zero-filled for -1 through cp.async
//   src-size 0), B tile = slice of the pre-transposed bf16 weights.  No scatter-add, no atomics,
//   fixed summation order.
//
// wgrad kernel: d_weight[k] = A_k^T B_k over the pair list of offset k.  The gathered rows are
//   [pairs][channels] = MN-major operands for UMMA (the reduction index is the pair), so the same
//   swizzled row layout is used with the MN-major bits set in the instruction descriptor.
//   CTAs split the pair list; partial tiles are reduced with fp32 atomics.
//
// Activations arrive as fp32 [rows, C] from the torch modules between the convolutions; a small
// cast kernel makes the bf16 copy (row pitch padded to 8 elements so every row is 16-byte aligned)
// that the K gathers then read at half the bytes.
#include "common.cuh"
#include "umma.cuh"

namespace wfsp {
namespace {

using namespace umma;

constexpr int kProducerWarps = 16;  // four per scheduler; eight were measured equal on small launches and slower on the large ones
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 32;  // + the MMA issuer warp
constexpr int kMmaWarp = kProducerWarps;
constexpr int kWeightWarp = kMmaWarp + 1;          // conv_apply only: issues the bulk copy of every weight slice
constexpr int kApplyThreads = kThreads + 32;
constexpr int kEpiWarps = 8;  // producer warps 0..7 drain TMEM (pairs w, w+4 share lane quarter w)
constexpr int kRowStep = kProducerThreads / 8;   // tile rows covered by one pass of the producers (8 threads per row)
constexpr int kRowsPerThread = 128 / kRowStep;   // rows of a 128-row tile per producer thread
constexpr int kTileM = 128;   // destination rows (apply) / a-channels (wgrad) per CTA
constexpr int kSliceK = 64;   // reduction elements per stage (one 128 B swizzle row of bf16)
constexpr int kABytes = kTileM * 128;
constexpr int kMaxStages = 8;
constexpr int kNbrStageK = 32;  // kernel volumes up to this keep the CTA's neighbour tile in smem
constexpr int kSmemMax = 222 * 1024;
// The opt-in limit is set to the same (maximum) value at every launch: a CUDA graph replays kernel nodes
// with whatever the function attribute is at replay time, so per-launch values would let a later, smaller
// launch of the same function starve an earlier node.
constexpr int kSmemOptIn = 226 * 1024;  // 227 KB per CTA minus room for the kernels' static shared memory
constexpr int kMaxRowBlocks = 4;  // 128-row blocks per CTA of the apply kernel
constexpr int kEpiPitch = 36;  // words per row of a warp's 32 x 32 epilogue tile (16-byte aligned, conflict-free)

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive_noinc_u32(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) { cp_async_arrive_noinc_u32(smem_u32(bar)); }

// ---- weight preparation: fp32 [kvol][c_red][c_dst] (or transposed) -> bf16 tiles
// [kvol][kc_pad/64][n_pad][64], zero padded, each row's eight 16-byte chunks already permuted by the
// 128-byte swizzle (chunk ^ (row & 7)).  Rows n0 .. n0+n_tile of one (offset, 64-channel slice) are then
// one contiguous block of n_tile * 128 bytes that a single bulk copy (cp.async.bulk) drops into shared
// memory exactly as the UMMA descriptor expects it.
// split3 (WFSP_MATH_BF16X3): the reduction is three segments of the real channel count c_red / 3 holding
// hi(w), lo(w), hi(w) -- hi = w rounded to bf16, lo = (w - hi) rounded to bf16 -- against activations laid out as
// [hi(a) | hi(a) | lo(a)]: a w ~ hi hi + hi lo + lo hi with fp32 accumulation (error ~2^-16 per product).
__device__ __forceinline__ float prep_weight_value(const float* __restrict__ w, int64_t i, int c_red, int c_dst,
                                                   int transpose_w, int n_pad, int num_kb, int split3 = 0) {
  const int e = int(i & 7), pc = int((i >> 3) & 7);
  int64_t t = i >> 6;
  const int n = int(t % n_pad);
  t /= n_pad;
  const int kb = int(t % num_kb), k = int(t / num_kb);
  int c = kb * 64 + ((pc ^ (n & 7)) << 3) + e;
  if (c >= c_red || n >= c_dst) return 0.f;
  int seg = 0, cr = c_red;
  if (split3) { cr = c_red / 3; seg = c / cr; c -= seg * cr; }
  const float* wk = w + int64_t(k) * cr * c_dst;
  const float v = transpose_w ? wk[int64_t(n) * cr + c] : wk[int64_t(c) * c_dst + n];
  if (!split3) return v;
  const float hi = __bfloat162float(__float2bfloat16_rn(v));
  return seg == 1 ? v - hi : hi;
}

// ---- activation cast: fp32 [n][c] -> bf16 [n][c_pad], zero padded; one 16-byte chunk per thread
// mode: 0 = round to bf16; 1 = [hi | hi | lo] over 3c channels (c_pad = pitch of 3c); 2 = hi only (= 0); 3 = lo only
struct CastJob { const float* src; __nv_bfloat16* dst; int64_t n; int c, c_pad; const int32_t* n_dev; int mode; };

__device__ __forceinline__ float cast_value(const CastJob& j, const float* row, int col) {
  if (j.mode == 1) {
    const int seg = col / j.c, cc = col - seg * j.c;
    if (seg > 2) return 0.f;
    const float v = __ldg(row + cc), hi = __bfloat162float(__float2bfloat16_rn(v));
    return seg == 2 ? v - hi : hi;
  }
  if (col >= j.c) return 0.f;
  const float v = __ldg(row + col);
  return j.mode == 3 ? v - __bfloat162float(__float2bfloat16_rn(v)) : v;
}

__global__ void __launch_bounds__(256) cast_rows_kernel(CastJob j0, CastJob j1) {
  const int64_t n0 = j0.n_dev ? int64_t(*j0.n_dev) : j0.n;
  const int64_t n1 = j1.n_dev ? int64_t(*j1.n_dev) : j1.n;
  const int64_t chunks0 = n0 * (j0.c_pad >> 3), chunks_total = chunks0 + n1 * (j1.c_pad >> 3);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < chunks_total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const CastJob& j = i < chunks0 ? j0 : j1;
    const int64_t ii = i < chunks0 ? i : i - chunks0;
    const int cpr = j.c_pad >> 3;
    const int64_t row = ii / cpr;
    const int col = int(ii - row * cpr) << 3;
    const float* s = j.src + row * j.c;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = cast_value(j, s, col + e);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(j.dst + row * j.c_pad + col) = u;
  }
}

// weight preparation and activation cast of one conv_apply call in a single launch: the first
// `prep_blocks` CTAs transpose / pad the weights, the rest cast the activation rows
__global__ void __launch_bounds__(256) prep_and_cast_kernel(const float* __restrict__ w, int kvol, int c_red, int c_dst,
                                                            int transpose_w, __nv_bfloat16* __restrict__ wt, int n_pad,
                                                            int kc_pad, int prep_blocks, CastJob job, int split3) {
  if (int(blockIdx.x) < prep_blocks) {
    const int num_kb = kc_pad / 64;
    const int64_t total = int64_t(kvol) * n_pad * kc_pad;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(prep_blocks) * blockDim.x)
      wt[i] = __float2bfloat16_rn(prep_weight_value(w, i, c_red, c_dst, transpose_w, n_pad, num_kb, split3));
    return;
  }
  const int64_t n = job.n_dev ? int64_t(*job.n_dev) : job.n;
  const int cpr = job.c_pad >> 3;
  const int64_t chunks = n * cpr;
  const int64_t nb = int64_t(gridDim.x) - prep_blocks;
  for (int64_t i = (int64_t(blockIdx.x) - prep_blocks) * blockDim.x + threadIdx.x; i < chunks; i += nb * blockDim.x) {
    const int64_t row = i / cpr;
    const int col = int(i - row * cpr) << 3;
    const float* s = job.src + row * job.c;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = cast_value(job, s, col + e);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(job.dst + row * job.c_pad + col) = u;
  }
}

// weight preparation of several layers / directions in one launch (the weights of a training step are
// fixed until the optimiser runs, so every layout the step needs can be made up front)
constexpr int kMaxPrepJobs = 16;
struct PrepJob { const float* w; __nv_bfloat16* out; int kvol, c_red, c_dst, transpose_w; int block0, blocks; };
struct PrepBatch { PrepJob job[kMaxPrepJobs]; int n; };

// One 16-byte chunk (eight consecutive reduction channels of one output channel) per thread and iteration: one
// index decomposition and one store per eight values.  The grid is kept BELOW one CTA per SM over all jobs: the launch
// runs beside the first convolution of the step, whose CTAs take a whole SM's shared memory (226 KB + the 1 KB every
// resident CTA reserves) -- a preparation launch that touched every SM kept them out until it had drained.
__global__ void __launch_bounds__(256) prep_weights_batch_kernel(const PrepBatch b) {
  int j = 0;
  while (j + 1 < b.n && int(blockIdx.x) >= b.job[j + 1].block0) ++j;
  const PrepJob& q = b.job[j];
  const int n_pad = (q.c_dst + 15) / 16 * 16, kc_pad = (q.c_red + 63) / 64 * 64, num_kb = kc_pad / 64;
  const int64_t chunks = int64_t(q.kvol) * n_pad * (kc_pad >> 3);
  uint4* out = reinterpret_cast<uint4*>(q.out);
  for (int64_t ci = (int64_t(blockIdx.x) - q.block0) * blockDim.x + threadIdx.x; ci < chunks; ci += int64_t(q.blocks) * blockDim.x) {
    const int pc = int(ci & 7);
    int64_t t = ci >> 3;
    const int n = int(t % n_pad);
    t /= n_pad;
    const int kb = int(t % num_kb), k = int(t / num_kb);
    const int c0 = kb * 64 + ((pc ^ (n & 7)) << 3);
    const float* wk = q.w + int64_t(k) * q.c_red * q.c_dst;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      v[e] = (c < q.c_red && n < q.c_dst) ? (q.transpose_w ? wk[int64_t(n) * q.c_red + c] : wk[int64_t(c) * q.c_dst + n]) : 0.f;
    }
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    out[ci] = u;
  }
}

int launch_cast(const CastJob& a, const CastJob* b, cudaStream_t st) {
  CastJob j1 = b ? *b : CastJob{nullptr, nullptr, 0, 8, 8, nullptr, 0};
  const int64_t c0 = a.n * (a.c_pad >> 3), c1 = j1.n * (j1.c_pad >> 3);
  if (c0 + c1 == 0) return WFSP_OK;
  int64_t blocks = ceil_div<int64_t>(c0 + c1, 256);
  if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
  cast_rows_kernel<<<unsigned(blocks), 256, 0, st>>>(a, j1);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

struct PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t free_[kMaxStages];
  uint64_t done;
};

__device__ __forceinline__ void init_pipe(PipeBarriers& b, uint32_t full_count) {
  for (int s = 0; s < kMaxStages; ++s) {
    mbar_init(&b.full[s], full_count);
    mbar_init(&b.free_[s], 1);
  }
  mbar_init(&b.done, 1);
  fence_mbar_init();
}

// one arrival on `bar` that also announces `bytes` of bulk-copy traffic to wait for
__device__ __forceinline__ void mbar_arrive_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// TMA bulk copy (no tensor map): `bytes` contiguous bytes global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s_u32(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}


// TMEM -> registers (thread = row) -> the warp's shared-memory tile -> global, the warp writing whole row
// segments so the stores are coalesced whatever the channel count (float4 / float2 / scalar by the
// alignment the row pitch allows).  acc: 32 columns of this thread's row.
template <int V>
__device__ __forceinline__ void store_tile_rows(const float* tile, float* dst, int64_t pitch, int rmax, int ncols,
                                                int c_left, const float* bias, int lane) {
  constexpr int kLanesPerRow = 32 / V;  // lanes covering the 32 columns of one row
  constexpr int kRowsPerIt = 32 / kLanesPerRow;
  const int c = (lane % kLanesPerRow) * V;
  if (c >= ncols || c >= c_left) return;  // V divides c_dst, so a vector is never cut by the row end
  float bv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + c + e) : 0.f;
#pragma unroll 4
  for (int r = lane / kLanesPerRow; r < rmax; r += kRowsPerIt) {
    float v[V];
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = tile[r * kEpiPitch + c + e] + bv[e];
    float* o = dst + int64_t(r) * pitch + c;
    if (V == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
    else o[0] = v[0];
  }
}

struct ApplyParams {
  const __nv_bfloat16* src; int64_t n_src; int c_pad;
  const __nv_bfloat16* wt; int n_pad, kc_pad;
  const float* bias; const int32_t* nbr; int kvol;
  float* dst; int64_t n_dst; int c_dst;
  int n_tile, stages, rblk, acc_stride, staged, nbr_bytes;
  const int32_t* n_src_dev; const int32_t* n_dst_dev;
  float* stats;  // optional [ceil(n_dst/32)][2][c_dst]: per 32-row chunk (mean, M2) of every output column
  // optional (dgrad): this output is the dy of a BatchNorm(+ReLU) whose input was bwd_x -- per 32-row chunk
  // [chunk][2][c_dst] sums of dy' and dy' * xhat (dy' = dy masked by the ReLU), so BatchNorm backward needs no
  // reduction pass over (x, dy)
  const float* bwd_x; const float* bwd_mean; const float* bwd_invstd; const float* bwd_gamma; const float* bwd_beta;
  float* bwd_part; int bwd_relu;
  int ksplit;  // CTAs of a thread-block cluster (along z) that split the (offset, channel slice) loop of one tile
  // optional (small launches, conv_apply_split_kernel): the BatchNorm(+ReLU, +Dropout) that follows this convolution,
  // finished INSIDE the launch -- after a grid-wide barrier every CTA folds the statistics partials of its columns and
  // normalises the blocks it produced (src/models/SPConvBlocks.py:505-510: conv . BatchNorm1d . ReLU . Dropout)
  int bn_fuse, bn_relu;
  const float* bn_gamma; const float* bn_beta;
  float* bn_rmean; float* bn_rvar; float* bn_save_mean; float* bn_save_invstd;
  float bn_eps, bn_momentum;
  float* bn_y32; __nv_bfloat16* bn_y16;
  unsigned* bn_bar;  // [2] {arrivals, generation}: zero before the first use, left consistent by every launch
  DropSpec bn_drop;
  unsigned long long* trace;  // debugging aid (wfsp_debug_trace): clock64 stamps of tile (0, 0)'s phases, 16 per cluster rank
};
#define WFSP_TRACE(slot)                                                                     \
  do {                                                                                       \
    if (p.trace != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)                \
      p.trace[krank * 16 + (slot)] = (unsigned long long)clock64();                          \
  } while (0)

// ---- thread-block cluster helpers (split reduction of small launches)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32x4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Grid-wide barrier of the `expected` live CTAs of a launch whose CTAs are all co-resident (the host checks the
// occupancy before it asks for a fused BatchNorm).  bar[0] counts arrivals and returns to zero, bar[1] is the
// generation the waiters watch.  Bounded spin: a launch that could not become co-resident traps instead of hanging.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = *reinterpret_cast<volatile unsigned*>(bar + 1);
    if (atomicAdd(bar, 1u) == expected - 1u) {
      atomicExch(bar, 0u);
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      const long long t0 = clock64();
      while (*reinterpret_cast<volatile unsigned*>(bar + 1) == gen) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// Phase 2 of a fused BatchNorm: one 32-row x (16|32)-column block of the convolution output (written by this warp
// before the grid barrier, so it comes back from L2) is normalised with the statistics folded from ALL chunks'
// partials -- the same sums in the same order as bn_apply's fused small path (csrc/bn.cu: stats_from_partials), so the
// result does not depend on which path ran -- and written as the next layer's bf16 operand (and / or fp32).
__device__ __forceinline__ void bn_finish_block(const ApplyParams& p, int64_t grow0, int rmax, int cc, int ncols, int lane,
                                                unsigned long long dkey) {
  if (rmax <= 0) return;
  const int c = p.c_dst, c_pad = (c + 7) & ~7;
  const int ccol = cc + lane;
  const bool on = lane < ncols && ccol < c;
  float mean = 0.f, invstd = 0.f, g = 1.f, bt = 0.f;
  if (on) {
    const int64_t n = p.n_dst, nblk = (n + 31) / 32;
    double a = 0.0, q = 0.0;
#pragma unroll 4
    for (int64_t b = 0; b < nblk; ++b) {
      const double nb = double(b + 1 < nblk ? 32 : n - b * 32);
      const double mb = __ldcg(p.stats + (b * 2 + 0) * c + ccol), m2b = __ldcg(p.stats + (b * 2 + 1) * c + ccol);
      a += nb * mb;
      q += m2b + nb * mb * mb;
    }
    const double cnt = double(n), mu = a / cnt;
    double m2 = q - cnt * mu * mu;
    if (m2 < 0.0) m2 = 0.0;
    mean = float(mu);
    invstd = float(1.0 / sqrt(m2 / cnt + double(p.bn_eps)));
    if (grow0 == 0) {  // exactly one block per column starts at row 0: it records the statistics
      p.bn_save_mean[ccol] = mean;
      p.bn_save_invstd[ccol] = invstd;
      if (p.bn_rmean) p.bn_rmean[ccol] = (1.f - p.bn_momentum) * p.bn_rmean[ccol] + p.bn_momentum * mean;
      if (p.bn_rvar && cnt > 1.0) p.bn_rvar[ccol] = (1.f - p.bn_momentum) * p.bn_rvar[ccol] + p.bn_momentum * float(m2 / (cnt - 1.0));
    }
    if (p.bn_gamma) g = __ldg(p.bn_gamma + ccol);
    if (p.bn_beta) bt = __ldg(p.bn_beta + ccol);
  }
  const bool pad = lane < ncols && ccol >= c && ccol < c_pad;  // zero padding of the bf16 operand rows
  const float* xp = p.dst + grow0 * c + ccol;
  for (int r0 = 0; r0 < rmax; r0 += 8) {
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = (on && r0 + u < rmax) ? __ldcg(xp + int64_t(r0 + u) * c) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u >= rmax) continue;
      const int64_t gr = grow0 + r0 + u;
      if (on) {
        float v = (xv[u] - mean) * invstd * g + bt;
        if (p.bn_relu && v < 0.f) v = 0.f;
        if (p.bn_drop.p > 0.f) v *= drop_factor(dkey, p.bn_drop.p, p.bn_drop.scale, (unsigned long long)(gr * c + ccol));
        if (p.bn_y32) p.bn_y32[gr * c + ccol] = v;
        if (p.bn_y16) p.bn_y16[gr * c_pad + ccol] = __float2bfloat16_rn(v);
      } else if (pad && p.bn_y16) {
        p.bn_y16[gr * c_pad + ccol] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// What happens to one finished 32-row x (16|32)-column block of the output, sitting in the warp's shared-memory
// tile: optional BatchNorm statistics of the chunk, optional BatchNorm-backward partial sums, then the coalesced
// store of the rows (+bias).
__device__ __forceinline__ void finish_block(const ApplyParams& p, const float* tile, int64_t grow0 /* first row */,
                                             int rmax, int cc /* first column */, int ncols, int vec, int lane) {
  if (rmax <= 0) return;
  if (p.stats != nullptr && lane < ncols && cc + lane < p.c_dst) {
    // BatchNorm statistics of this 32-row chunk, straight from the tile (saves the pass that would
    // re-read the output from HBM): shifted sums -> (mean, M2), merged later by bn_finalize_stats
    const int ccol = cc + lane;
    const float bv = p.bias ? __ldg(p.bias + ccol) : 0.f;
    const float K = tile[lane];
    float sd = 0.f, sq = 0.f;
    for (int r = 0; r < rmax; ++r) {
      const float d = tile[r * kEpiPitch + lane] - K;
      sd += d;
      sq += d * d;
    }
    const float cnt = float(rmax);
    float* sp = p.stats + (grow0 >> 5) * 2 * p.c_dst + ccol;
    sp[0] = K + bv + sd / cnt;
    sp[p.c_dst] = fmaxf(sq - sd * sd / cnt, 0.f);
  }
  if (p.bwd_part != nullptr && lane < ncols && cc + lane < p.c_dst) {
    const int ccol = cc + lane;
    const float m = __ldg(p.bwd_mean + ccol), is = __ldg(p.bwd_invstd + ccol);
    const float g = p.bwd_gamma ? __ldg(p.bwd_gamma + ccol) : 1.f, b = p.bwd_beta ? __ldg(p.bwd_beta + ccol) : 0.f;
    const float* xp = p.bwd_x + grow0 * p.c_dst + ccol;
    float s0 = 0.f, s1 = 0.f;
    for (int r0 = 0; r0 < rmax; r0 += 8) {
      float xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = (r0 + u < rmax) ? __ldg(xp + int64_t(r0 + u) * p.c_dst) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r0 + u < rmax) {
          const float xh = (xv[u] - m) * is;
          float d = tile[(r0 + u) * kEpiPitch + lane];
          if (p.bwd_relu && xh * g + b <= 0.f) d = 0.f;
          s0 += d;
          s1 += d * xh;
        }
      }
    }
    float* sp = p.bwd_part + (grow0 >> 5) * 2 * p.c_dst + ccol;
    sp[0] = s0;
    sp[p.c_dst] = s1;
  }
  float* o = p.dst + grow0 * p.c_dst + cc;
  const float* bias = p.bias ? p.bias + cc : nullptr;
  if (vec == 4) store_tile_rows<4>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
  else if (vec == 2) store_tile_rows<2>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
  else store_tile_rows<1>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
}

// A CTA owns rblk (1..4) consecutive 128-row blocks of destination rows and one column tile.  Every
// pipeline stage holds the gathered A slice of each row block plus ONE weight slice shared by all of
// them, so the weights -- the larger half of the operand traffic when the channel counts are a few
// hundred -- cross L2 -> SM once per rblk row blocks.  Accumulators: rblk x n_tile fp32 columns of TMEM.
template <int RB>
__global__ void __launch_bounds__(kApplyThreads) conv_apply_umma_kernel(const ApplyParams pp) {
  ApplyParams p = pp;
  if (p.n_src_dev) p.n_src = *p.n_src_dev;
  if (p.n_dst_dev) p.n_dst = *p.n_dst_dev;
  const int64_t row0 = int64_t(blockIdx.x) * (kTileM * RB);
  if (row0 >= p.n_dst) return;  // capacity-sized grid: nothing live in this tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ PipeBarriers bars;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_active[WFSP_MAX_KVOL / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.y * p.n_tile;
  int rows_left = int(p.n_dst - row0 < int64_t(kTileM * RB) ? p.n_dst - row0 : int64_t(kTileM * RB));
  const int rb_live = (rows_left + kTileM - 1) / kTileM;  // row blocks with at least one live row
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = uint32_t(RB) * kABytes;
  const uint32_t stage_bytes = a_bytes + uint32_t(p.n_tile) * 128u;
  int32_t* s_nbr = reinterpret_cast<int32_t*>(smem + uint32_t(p.stages) * stage_bytes);
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(RB * p.acc_stride));
  const bool staged = p.nbr != nullptr && p.staged;

  for (int i = tid; i < WFSP_MAX_KVOL / 32; i += kApplyThreads) s_active[i] = 0;
  if (tid == 0) init_pipe(bars, kProducerThreads + 1);
  if (warp == kMmaWarp) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // neighbour tile of this CTA: which kernel offsets are active, and (small kernels) a smem copy.
  // Loads are issued eight at a time before any is used: one memory round trip per batch.
  if (p.nbr) {
    const int total = rb_live * kTileM * p.kvol;
    const int limit = rows_left * p.kvol;
    const int32_t* base = p.nbr + row0 * p.kvol;
    for (int i0 = tid; i0 < total; i0 += kApplyThreads * 8) {
      int v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * kApplyThreads;
        v[j] = i < limit ? __ldg(base + i) : -1;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j * kApplyThreads;
        if (i < total) {
          const int vv = v[j] >= p.n_src ? -1 : v[j];
          if (staged) s_nbr[i] = vv;
          if (vv >= 0) {
            const int k = i % p.kvol;
            atomicOr(&s_active[k >> 5], 1u << (k & 31));
          }
        }
      }
    }
  } else if (tid == 0) {
    s_active[0] = 1u;
  }
  __syncthreads();

  const int num_kb = p.kc_pad / kSliceK;
  int n_active = 0;
  for (int w = 0; w < (p.kvol + 31) / 32; ++w) n_active += __popc(s_active[w]);
  const int total_iters = n_active * num_kb;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(&bars.full[0]), free0 = smem_u32(&bars.free_[0]);  // barrier s is 8 s bytes further

  if (warp < kProducerWarps) {
    // ------------------------------------------------------------------ producers
    // A: 16-byte cp.async of this thread's chunk of 8 rows per row block (rows rsub + 16 i are 2048 B
    // apart in the swizzled tile); B: thread 0 issues one bulk copy of the pre-swizzled weight slice.
    const int c16 = tid & 7;    // 16-byte chunk inside the 128-byte slice row
    const int rsub = tid >> 3;  // this thread covers tile rows rsub + kRowStep*i
    const uint32_t off0 = sw128_offset(uint32_t(rsub), uint32_t(c16));
    const int kb_lim = (p.c_pad - c16 * 8 + kSliceK - 1) / kSliceK;  // slices in which the chunk is inside the row
    const size_t a_row_bytes = size_t(p.c_pad) * 2;
    const char* src_c = reinterpret_cast<const char*>(p.src) + c16 * 16;
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int kw = 0; kw < (p.kvol + 31) / 32; ++kw) {
      uint32_t mask = s_active[kw];
      while (mask) {
        const int k = kw * 32 + __ffs(mask) - 1;
        mask &= mask - 1;
        int rows[RB * kRowsPerThread];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
#pragma unroll
          for (int i = 0; i < kRowsPerThread; ++i) {
            const int r = rb * kTileM + rsub + kRowStep * i;
            int v = -1;
            if (rb < rb_live) {
              if (!p.nbr) {
                v = r < rows_left ? int(row0 + r) : -1;
              } else if (staged) {
                v = s_nbr[r * p.kvol + k];
              } else if (r < rows_left) {
                v = __ldg(p.nbr + (row0 + r) * p.kvol + k);
                if (v >= p.n_src) v = -1;
              }
            }
            rows[rb * kRowsPerThread + i] = v;
          }
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          if (it >= p.stages) mbar_wait_u32(free0 + 8u * s, ph ^ 1u);
          const uint32_t sa = smem0 + uint32_t(s) * stage_bytes;
          const bool col_ok = kb < kb_lim;
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
            if (rb < rb_live) {
#pragma unroll
              for (int i = 0; i < kRowsPerThread; ++i) {
                const int v = rows[rb * kRowsPerThread + i];
                const bool ok = col_ok && v >= 0;
                cp_async16(sa + off0 + rb * kABytes + i * (kRowStep * 128), src_c + (ok ? size_t(v) * a_row_bytes + kb * 128 : size_t(0)),
                           ok ? 16u : 0u);
              }
            }
          }
          cp_async_arrive_noinc_u32(full0 + 8u * s);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    // ------------------------------------------------------------------ epilogue
    // The pipeline's stage memory is free once `done` has fired: warp w stages its 32 x 32 blocks there.
    if (total_iters > 0) {
      mbar_wait(&bars.done, 0);
      tc_fence_after();
    }
    float* tile = reinterpret_cast<float*>(smem) + (warp % kEpiWarps) * (32 * kEpiPitch);
    const bool al16 = (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0, al8 = (reinterpret_cast<uintptr_t>(p.dst) & 7) == 0;
    const int vec = ((p.c_dst & 3) == 0 && al16) ? 4 : (((p.c_dst & 1) == 0 && al8) ? 2 : 1);
    for (int rb = 0; rb < (warp < kEpiWarps ? rb_live : 0); ++rb) {
      // warps w and w+4 share TMEM lane quarter w (a warp may only read lanes 32*(warp%4)..+31): they take
      // alternate 32-column chunks of it
      const int wr0 = rb * kTileM + (warp & 3) * 32;  // first row of this warp's block inside the CTA tile
      int rmax = rows_left - wr0;
      if (rmax > 32) rmax = 32;
      for (int col = (warp >> 2) * 32; col < p.n_tile; col += 64) {
        const int ncols = p.n_tile - col < 32 ? p.n_tile - col : 32;  // 16 or 32
        uint32_t acc[32];
        if (total_iters > 0) {
          const uint32_t taddr = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t(rb * p.acc_stride + col);
          tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
          if (ncols > 16) tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (4 * q < ncols)
            *reinterpret_cast<uint4*>(tile + lane * kEpiPitch + 4 * q) = make_uint4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        __syncwarp();
        if (rmax > 0 && p.stats != nullptr && lane < ncols && n0 + col + lane < p.c_dst) {
          // BatchNorm statistics of this 32-row chunk, straight from the tile (saves the pass that would
          // re-read the output from HBM): shifted sums -> (mean, M2), merged later by bn_finalize_stats
          const int ccol = n0 + col + lane;
          const float bv = p.bias ? __ldg(p.bias + ccol) : 0.f;
          const float K = tile[lane];
          float sd = 0.f, sq = 0.f;
          for (int r = 0; r < rmax; ++r) {
            const float d = tile[r * kEpiPitch + lane] - K;
            sd += d;
            sq += d * d;
          }
          const float cnt = float(rmax);
          float* sp = p.stats + ((row0 + wr0) >> 5) * 2 * p.c_dst + ccol;
          sp[0] = K + bv + sd / cnt;
          sp[p.c_dst] = fmaxf(sq - sd * sd / cnt, 0.f);
        }
        if (rmax > 0) {
          const int cc = n0 + col;
          float* o = p.dst + (row0 + wr0) * p.c_dst + cc;
          const float* bias = p.bias ? p.bias + cc : nullptr;
          if (vec == 4) store_tile_rows<4>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
          else if (vec == 2) store_tile_rows<2>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
          else store_tile_rows<1>(tile, o, p.c_dst, rmax, ncols, p.c_dst - cc, bias, lane);
        }
        __syncwarp();
      }
    }
  } else if (warp == kWeightWarp) {
    // ------------------------------------------------------------------ weight slices
    // One bulk copy (cp.async.bulk) per pipeline stage drops the pre-swizzled [n_tile][64] slice of the current
    // (offset, 64-channel block) behind the gathered rows.  A warp of its own: the copy's operands live on the
    // uniform datapath, which a lane of a (divergent) gather warp can only reach through register moves.
    int n_lim = p.n_pad - n0;  // the last column tile may overhang the padded weights
    if (n_lim > p.n_tile) n_lim = p.n_tile;
    const uint32_t b_bytes = uint32_t(n_lim) * 128u;
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int kw = 0; kw < (p.kvol + 31) / 32; ++kw) {
      uint32_t mask = s_active[kw];
      while (mask) {
        const int k = kw * 32 + __ffs(mask) - 1;
        mask &= mask - 1;
        const char* wk = reinterpret_cast<const char*>(p.wt) + ((size_t(k) * num_kb) * p.n_pad + n0) * 128;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          if (it >= p.stages) mbar_wait_u32(free0 + 8u * s, ph ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx_u32(full0 + 8u * s, b_bytes);
            bulk_copy_g2s_u32(smem0 + uint32_t(s) * stage_bytes + a_bytes, wk + size_t(kb) * p.n_pad * 128, b_bytes, full0 + 8u * s);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the pipeline (converged control flow keeps the descriptor arithmetic on the uniform
    // datapath); one elected lane issues.  The descriptors of the four K=16 steps of a slice differ only in
    // their start-address field: +32 bytes = +2 in the low word.
    const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 0, 0);
    const uint64_t desc_hi = make_desc_sw128(0, 16, 1024);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < total_iters; ++it) {
      mbar_wait_u32(full0 + 8u * s, ph);
      fence_proxy_async_smem();
      tc_fence_after();
      const uint32_t a_addr = smem0 + uint32_t(s) * stage_bytes, b_addr = a_addr + a_bytes;
      if (elect_one()) {
        const uint64_t bdesc0 = desc_hi | uint64_t((b_addr >> 4) & 0x3fffu);
        for (int rb = 0; rb < rb_live; ++rb) {
          const uint64_t adesc0 = desc_hi | uint64_t(((a_addr + rb * kABytes) >> 4) & 0x3fffu);
#pragma unroll
          for (int kk = 0; kk < kSliceK / 16; ++kk)
            mma_bf16(tmem + uint32_t(rb * p.acc_stride), adesc0 + 2 * kk, bdesc0 + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        mma_commit_u32(free0 + 8u * s);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (total_iters > 0 && elect_one()) mma_commit(&bars.done);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, tmem_cols);
}

// Small launches: the same tile walk, but the (offset, slice) loop of a tile is split over the CTAs of a thread-block
// cluster (SPLIT), the grid covers only the EXPECTED live rows and the kernel loops over tiles.  Kept apart from the
// throughput kernel above: under its 96-register budget any extra live value costs the hot loops there.
template <int RB, bool SPLIT>
__global__ void __launch_bounds__(kApplyThreads) conv_apply_split_kernel(const ApplyParams pp) {
  ApplyParams p = pp;
  if (p.n_src_dev) p.n_src = *p.n_src_dev;
  if (p.n_dst_dev) p.n_dst = *p.n_dst_dev;
  constexpr int kTileRows = kTileM * RB;
  if (int64_t(blockIdx.x) * kTileRows >= p.n_dst) return;  // nothing live for this CTA (uniform over a cluster: same blockIdx.x)
  extern __shared__ uint8_t smem_raw[];
  __shared__ PipeBarriers bars;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_active[WFSP_MAX_KVOL / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.y * p.n_tile;
  const int ks = SPLIT ? p.ksplit : 1, krank = SPLIT ? int(blockIdx.z) : 0;  // cluster (1, 1, ks): rank = blockIdx.z
  WFSP_TRACE(0);
  if (p.trace != nullptr && tid == 0) {  // %globaltimer at entry: slot 15 of tile (0, 0)'s ranks, and every CTA's from slot 128
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    if (blockIdx.x == 0 && blockIdx.y == 0) p.trace[krank * 16 + 15] = gt;
    const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (lin < 1024) p.trace[128 + 2 * lin] = gt;
  }
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = uint32_t(RB) * kABytes;
  const uint32_t stage_bytes = a_bytes + uint32_t(p.n_tile) * 128u;
  int32_t* s_nbr = reinterpret_cast<int32_t*>(smem + uint32_t(p.stages) * stage_bytes);
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(RB * p.acc_stride));
  const bool staged = p.nbr != nullptr && p.staged;
  const int num_kb = p.kc_pad / kSliceK;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(&bars.full[0]), free0 = smem_u32(&bars.free_[0]);  // barrier s is 8 s bytes further
  const bool al16 = (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0, al8 = (reinterpret_cast<uintptr_t>(p.dst) & 7) == 0;
  const int vec = ((p.c_dst & 3) == 0 && al16) ? 4 : (((p.c_dst & 1) == 0 && al8) ? 2 : 1);
  const int ncc = (p.n_tile + 31) / 32;  // 32-column chunks of the tile (the last may be 16 wide)
  constexpr int kBlockFloats = 32 * kEpiPitch;
  // split launches: the blocks other CTAs of the cluster push here live behind the neighbour tile (never aliased
  // with the pipeline stages: a fast peer pushes while this CTA is still in its main loop)
  float* park = reinterpret_cast<float*>(smem + uint32_t(p.stages) * stage_bytes + uint32_t(p.nbr_bytes));

  if (warp == kMmaWarp) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }

  // The grid covers the EXPECTED live rows (launch-shape hint); whatever lies beyond is taken by further trips of
  // this loop, so a wrong hint costs time, never rows.
  for (int64_t tile_i = blockIdx.x; tile_i * kTileRows < p.n_dst; tile_i += gridDim.x) {
    const int64_t row0 = tile_i * kTileRows;
    const int rows_left = int(p.n_dst - row0 < int64_t(kTileRows) ? p.n_dst - row0 : int64_t(kTileRows));
    const int rb_live = (rows_left + kTileM - 1) / kTileM;  // row blocks with at least one live row
    for (int i = tid; i < WFSP_MAX_KVOL / 32; i += kApplyThreads) s_active[i] = 0;
    if (tid == 0) init_pipe(bars, kProducerThreads + 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    WFSP_TRACE(1);

    // neighbour tile of this CTA: which kernel offsets are active, and (small kernels) a smem copy.
    // Loads are issued eight at a time before any is used: one memory round trip per batch.
    if (p.nbr) {
      const int total = rb_live * kTileM * p.kvol;
      const int limit = rows_left * p.kvol;
      const int32_t* base = p.nbr + row0 * p.kvol;
      uint32_t mine = 0;  // kernel volumes up to 32: the warp ORs its bits and ONE lane updates the shared mask
      for (int i0 = tid; i0 < total; i0 += kApplyThreads * 8) {
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = i0 + j * kApplyThreads;
          v[j] = i < limit ? __ldg(base + i) : -1;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = i0 + j * kApplyThreads;
          if (i < total) {
            const int vv = v[j] >= p.n_src ? -1 : v[j];
            if (staged) s_nbr[i] = vv;
            if (vv >= 0) {
              const int k = i % p.kvol;
              if (p.kvol <= 32) mine |= 1u << k;
              else atomicOr(&s_active[k >> 5], 1u << (k & 31));
            }
          }
        }
      }
      if (p.kvol <= 32) {
        mine = __reduce_or_sync(0xffffffffu, mine);
        if (lane == 0 && mine) atomicOr(&s_active[0], mine);
      }
    } else if (tid == 0) {
      s_active[0] = 1u;
    }
    __syncthreads();

    int n_active = 0;
    for (int w = 0; w < (p.kvol + 31) / 32; ++w) n_active += __popc(s_active[w]);
    const int total_iters = n_active * num_kb;
    // this CTA's share of the flattened (active offset, 64-channel slice) loop; every CTA of a cluster sees the same
    // neighbour tile, hence the same total
    const int it_lo = SPLIT ? total_iters * krank / ks : 0;
    const int it_hi = SPLIT ? total_iters * (krank + 1) / ks : total_iters;
    const int my_iters = it_hi - it_lo;
    WFSP_TRACE(2);

    if (warp < kProducerWarps) {
      // ------------------------------------------------------------------ producers
      // A: 16-byte cp.async of this thread's chunk of 8 rows per row block (rows rsub + 16 i are 2048 B
      // apart in the swizzled tile); B: the weight warp issues one bulk copy of the pre-swizzled weight slice.
      const int c16 = tid & 7;    // 16-byte chunk inside the 128-byte slice row
      const int rsub = tid >> 3;  // this thread covers tile rows rsub + kRowStep*i
      const uint32_t off0 = sw128_offset(uint32_t(rsub), uint32_t(c16));
      const int kb_lim = (p.c_pad - c16 * 8 + kSliceK - 1) / kSliceK;  // slices in which the chunk is inside the row
      const size_t a_row_bytes = size_t(p.c_pad) * 2;
      const char* src_c = reinterpret_cast<const char*>(p.src) + c16 * 16;
      int s = 0, it = 0, g = 0;
      uint32_t ph = 0;
      for (int kw = 0; kw < (p.kvol + 31) / 32 && g < it_hi; ++kw) {
        uint32_t mask = s_active[kw];
        while (mask && g < it_hi) {
          const int k = kw * 32 + __ffs(mask) - 1;
          mask &= mask - 1;
          if (g + num_kb <= it_lo) { g += num_kb; continue; }  // another CTA of the cluster takes this offset
          int rows[RB * kRowsPerThread];
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i) {
              const int r = rb * kTileM + rsub + kRowStep * i;
              int v = -1;
              if (rb < rb_live) {
                if (!p.nbr) {
                  v = r < rows_left ? int(row0 + r) : -1;
                } else if (staged) {
                  v = s_nbr[r * p.kvol + k];
                } else if (r < rows_left) {
                  v = __ldg(p.nbr + (row0 + r) * p.kvol + k);
                  if (v >= p.n_src) v = -1;
                }
              }
              rows[rb * kRowsPerThread + i] = v;
            }
          }
          const int kb0 = g < it_lo ? it_lo - g : 0;
          const int kb1 = g + num_kb > it_hi ? it_hi - g : num_kb;
          g += num_kb;
          for (int kb = kb0; kb < kb1; ++kb) {
            if (it >= p.stages) mbar_wait_u32(free0 + 8u * s, ph ^ 1u);
            const uint32_t sa = smem0 + uint32_t(s) * stage_bytes;
            const bool col_ok = kb < kb_lim;
#pragma unroll
            for (int rb = 0; rb < RB; ++rb) {
              if (rb < rb_live) {
#pragma unroll
                for (int i = 0; i < kRowsPerThread; ++i) {
                  const int v = rows[rb * kRowsPerThread + i];
                  const bool ok = col_ok && v >= 0;
                  cp_async16(sa + off0 + rb * kABytes + i * (kRowStep * 128), src_c + (ok ? size_t(v) * a_row_bytes + kb * 128 : size_t(0)),
                             ok ? 16u : 0u);
                }
              }
            }
            cp_async_arrive_noinc_u32(full0 + 8u * s);
            ++it;
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
      // The pipeline's stage memory is free once `done` has fired: the epilogue stages its blocks there.
      WFSP_TRACE(3);
      if (my_iters > 0) {
        mbar_wait(&bars.done, 0);
        tc_fence_after();
      }
      WFSP_TRACE(4);
    } else if (warp == kWeightWarp) {
      // ------------------------------------------------------------------ weight slices
      // One bulk copy (cp.async.bulk) per pipeline stage drops the pre-swizzled [n_tile][64] slice of the current
      // (offset, 64-channel block) behind the gathered rows.  A warp of its own: the copy's operands live on the
      // uniform datapath, which a lane of a (divergent) gather warp can only reach through register moves.
      int n_lim = p.n_pad - n0;  // the last column tile may overhang the padded weights
      if (n_lim > p.n_tile) n_lim = p.n_tile;
      const uint32_t b_bytes = uint32_t(n_lim) * 128u;
      int s = 0, it = 0, g = 0;
      uint32_t ph = 0;
      for (int kw = 0; kw < (p.kvol + 31) / 32 && g < it_hi; ++kw) {
        uint32_t mask = s_active[kw];
        while (mask && g < it_hi) {
          const int k = kw * 32 + __ffs(mask) - 1;
          mask &= mask - 1;
          if (g + num_kb <= it_lo) { g += num_kb; continue; }
          const char* wk = reinterpret_cast<const char*>(p.wt) + ((size_t(k) * num_kb) * p.n_pad + n0) * 128;
          const int kb0 = g < it_lo ? it_lo - g : 0;
          const int kb1 = g + num_kb > it_hi ? it_hi - g : num_kb;
          g += num_kb;
          for (int kb = kb0; kb < kb1; ++kb) {
            if (it >= p.stages) mbar_wait_u32(free0 + 8u * s, ph ^ 1u);
            if (elect_one()) {
              mbar_arrive_expect_tx_u32(full0 + 8u * s, b_bytes);
              bulk_copy_g2s_u32(smem0 + uint32_t(s) * stage_bytes + a_bytes, wk + size_t(kb) * p.n_pad * 128, b_bytes, full0 + 8u * s);
            }
            __syncwarp();
            ++it;
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else {
      // ------------------------------------------------------------------ MMA issuer
      // The whole warp walks the pipeline (converged control flow keeps the descriptor arithmetic on the uniform
      // datapath); one elected lane issues.  The descriptors of the four K=16 steps of a slice differ only in
      // their start-address field: +32 bytes = +2 in the low word.
      const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 0, 0);
      const uint64_t desc_hi = make_desc_sw128(0, 16, 1024);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_iters; ++it) {
        mbar_wait_u32(full0 + 8u * s, ph);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t a_addr = smem0 + uint32_t(s) * stage_bytes, b_addr = a_addr + a_bytes;
        if (elect_one()) {
          const uint64_t bdesc0 = desc_hi | uint64_t((b_addr >> 4) & 0x3fffu);
          for (int rb = 0; rb < rb_live; ++rb) {
            const uint64_t adesc0 = desc_hi | uint64_t(((a_addr + rb * kABytes) >> 4) & 0x3fffu);
#pragma unroll
            for (int kk = 0; kk < kSliceK / 16; ++kk)
              mma_bf16(tmem + uint32_t(rb * p.acc_stride), adesc0 + 2 * kk, bdesc0 + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
          mma_commit_u32(free0 + 8u * s);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (my_iters > 0 && elect_one()) mma_commit(&bars.done);
    }

    // ------------------------------------------------------------------ epilogue (warps 0..7)
    // TMEM -> registers (thread = row) -> a 32 x 32 block in shared memory.  Warps w and w+4 share TMEM lane
    // quarter w (a warp may only read lanes 32*(warp%4)..+31): they take alternate 32-column chunks of it.
    auto load_acc = [&](int rb, int col, int ncols, uint32_t (&acc)[32]) {  // this warp's lane quarter of row block rb
      if (my_iters > 0) {
        const uint32_t taddr = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t(rb * p.acc_stride + col);
        tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        if (ncols > 16) tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = 0u;
      }
    };
    float* tile = reinterpret_cast<float*>(smem) + (warp % kEpiWarps) * kBlockFloats;
    if (!SPLIT) {
      if (warp < kEpiWarps) {
        for (int rb = 0; rb < rb_live; ++rb) {
          const int wr0 = rb * kTileM + (warp & 3) * 32;  // first row of this warp's block inside the CTA tile
          int rmax = rows_left - wr0;
          if (rmax > 32) rmax = 32;
          for (int col = (warp >> 2) * 32; col < p.n_tile; col += 64) {
            const int ncols = p.n_tile - col < 32 ? p.n_tile - col : 32;  // 16 or 32
            uint32_t acc[32];
            load_acc(rb, col, ncols, acc);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (4 * q < ncols)
                *reinterpret_cast<uint4*>(tile + lane * kEpiPitch + 4 * q) = make_uint4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            __syncwarp();
            finish_block(p, tile, row0 + wr0, rmax, n0 + col, ncols, vec, lane);
            __syncwarp();
          }
        }
      }
    } else {
      // Split reduction.  The tile's 32 x 32 blocks are dealt out round-robin over the CTAs of the cluster; every
      // CTA PUSHES its partial of a block into a slot of the owner's shared memory (remote stores do not stall),
      // one cluster barrier, then the owner adds the ks slots IN RANK ORDER (fixed summation order, no atomics)
      // from its own shared memory and finishes the block as above.
      if (warp < kEpiWarps) {
        for (int rb = 0; rb < rb_live; ++rb)
          for (int cc = warp >> 2; cc < ncc; cc += 2) {
            const int ncols = p.n_tile - cc * 32 < 32 ? p.n_tile - cc * 32 : 32;
            uint32_t acc[32];
            load_acc(rb, cc * 32, ncols, acc);
            const int u = (rb * 4 + (warp & 3)) * ncc + cc;
            const uint32_t slot = smem_u32(park + ((u / ks) * ks + krank) * kBlockFloats + lane * kEpiPitch);
            const uint32_t remote = map_to_cta(slot, uint32_t(u % ks));
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (4 * q < ncols) st_cluster_u32x4(remote + 16u * q, acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
          }
      }
      WFSP_TRACE(5);
      __syncwarp();
      cluster_sync_all();
      WFSP_TRACE(6);
      if (warp < kEpiWarps) {
        const int units = rb_live * 4 * ncc;
        for (int j = warp; krank + ks * j < units; j += kEpiWarps) {
          const int u = krank + ks * j;
          const int cc = u % ncc, rq = (u / ncc) & 3, rb = u / (4 * ncc);
          const int ncols = p.n_tile - cc * 32 < 32 ? p.n_tile - cc * 32 : 32;
          const int wr0 = rb * kTileM + rq * 32;
          int rmax = rows_left - wr0;
          if (rmax > 32) rmax = 32;
          if (rmax > 0) {
            // lane -> rows (lane >> 3) + 4 i, four columns from 4 (lane & 7)
            const int c4 = (lane & 7) * 4;
            const float* slot0 = park + (j * ks) * kBlockFloats + (lane >> 3) * kEpiPitch + c4;
            if (c4 < ncols) {
#pragma unroll 1
              for (int h = 0; h < 2; ++h) {  // two batches of four 16-byte loads per slot (register budget)
                const float* sl = slot0 + h * 16 * kEpiPitch;
                float4 acc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = *reinterpret_cast<const float4*>(sl + i * 4 * kEpiPitch);
                for (int src = 1; src < ks; ++src) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(sl + src * kBlockFloats + i * 4 * kEpiPitch);
                    acc[i].x += t.x; acc[i].y += t.y; acc[i].z += t.z; acc[i].w += t.w;
                  }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  *reinterpret_cast<float4*>(tile + ((lane >> 3) + 16 * h + 4 * i) * kEpiPitch + c4) = acc[i];
              }
            }
            __syncwarp();
            finish_block(p, tile, row0 + wr0, rmax, n0 + cc * 32, ncols, vec, lane);
            __syncwarp();
          }
        }
      }
      WFSP_TRACE(7);
    }
    WFSP_TRACE(8);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    // a further trip (wrong launch-shape hint only): peers must have summed this trip's slots before anyone pushes again
    if (SPLIT && (tile_i + gridDim.x) * kTileRows < p.n_dst) cluster_sync_all();
  }
  if (warp == kMmaWarp) tmem_dealloc(s_tmem, tmem_cols);
  if (p.bn_fuse) {
    // ---- fused BatchNorm(+ReLU, +Dropout): every output block and every statistics partial of the layer is in
    // memory once ALL live CTAs have passed this barrier; each warp then normalises the blocks it produced
    const int64_t tiles = (p.n_dst + kTileRows - 1) / kTileRows;
    const unsigned live_x = unsigned(tiles < int64_t(gridDim.x) ? tiles : int64_t(gridDim.x));
    grid_barrier(p.bn_bar, live_x * gridDim.y * gridDim.z);
    if (warp < kEpiWarps) {
      const unsigned long long dkey = p.bn_drop.p > 0.f ? drop_key(p.bn_drop) : 0ull;
      for (int64_t tile_i = blockIdx.x; tile_i * kTileRows < p.n_dst; tile_i += gridDim.x) {
        const int64_t row0 = tile_i * kTileRows;
        const int rows_left = int(p.n_dst - row0 < int64_t(kTileRows) ? p.n_dst - row0 : int64_t(kTileRows));
        const int rb_live = (rows_left + kTileM - 1) / kTileM;
        if (!SPLIT) {
          for (int rb = 0; rb < rb_live; ++rb) {
            const int wr0 = rb * kTileM + (warp & 3) * 32;
            int rmax = rows_left - wr0;
            if (rmax > 32) rmax = 32;
            for (int col = (warp >> 2) * 32; col < p.n_tile; col += 64)
              bn_finish_block(p, row0 + wr0, rmax, n0 + col, p.n_tile - col < 32 ? p.n_tile - col : 32, lane, dkey);
          }
        } else {
          const int units = rb_live * 4 * ncc;
          for (int j = warp; krank + ks * j < units; j += kEpiWarps) {
            const int u = krank + ks * j;
            const int cc = u % ncc, rq = (u / ncc) & 3, rb = u / (4 * ncc);
            const int wr0 = rb * kTileM + rq * 32;
            int rmax = rows_left - wr0;
            if (rmax > 32) rmax = 32;
            bn_finish_block(p, row0 + wr0, rmax, n0 + cc * 32, p.n_tile - cc * 32 < 32 ? p.n_tile - cc * 32 : 32, lane, dkey);
          }
        }
      }
    }
  }
  WFSP_TRACE(9);
  if (p.trace != nullptr && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (lin < 1024) p.trace[128 + 2 * lin + 1] = gt;
  }
}

struct WgradParams {
  const __nv_bfloat16* a; int64_t n_a; int c_a, ca_pad;
  const __nv_bfloat16* b; int64_t n_b; int c_b, cb_pad;
  const int32_t* pair_a; const int32_t* pair_b; const int32_t* pair_num;
  int kvol; int64_t pitch;
  float* dw;
  int n_tile, m_groups, nsplit, stages, use_atomic, acc_stride;
  int64_t chunk;
  const int32_t* n_a_dev;
};

constexpr int kPairsPerThread = kSliceK / kRowStep;  // pairs of a 64-pair slice per producer thread
constexpr int kIdxGroup = 4;  // pair-index loads are issued this many pipeline slices ahead, as one batch

// MT = 128-channel blocks of `a` per CTA: they share every gathered slice of `b` rows and its pair
// indices (the a-side of d_weight is at most a few hundred channels, so MT = 2 usually covers it).
template <int MT>
__global__ void __launch_bounds__(kThreads) conv_wgrad_umma_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ PipeBarriers bars;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.y / p.nsplit, split = blockIdx.y % p.nsplit;
  const int mg = blockIdx.x % p.m_groups, nt = blockIdx.x / p.m_groups;
  const int a_c0 = mg * (kTileM * MT), b_c0 = nt * p.n_tile;
  const int64_t n_pairs = p.pair_num ? int64_t(p.pair_num[k]) : (p.n_a_dev ? int64_t(*p.n_a_dev) : p.n_a);
  const int64_t begin = int64_t(split) * p.chunk;
  int64_t end = begin + p.chunk;
  if (end > n_pairs || split == p.nsplit - 1) end = n_pairs;  // the last split takes whatever the launch-shape hint missed
  if (begin >= n_pairs && (split > 0 || p.use_atomic)) return;  // nothing to add (uniform per CTA)
  const int iters = begin < end ? int((end - begin + kSliceK - 1) / kSliceK) : 0;
  int mt_live = (p.c_a - a_c0 + kTileM - 1) / kTileM;  // 128-channel blocks of this CTA that hold real channels
  if (mt_live > MT) mt_live = MT;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_panels = (p.n_tile + 63) / 64;
  constexpr uint32_t a_bytes = 2u * MT * 8192u;
  const uint32_t stage_bytes = a_bytes + uint32_t(b_panels) * 8192u;
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(MT * p.acc_stride));

  if (tid == 0) init_pipe(bars, kProducerThreads);
  if (warp == kMmaWarp) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t smem0 = smem_u32(smem);

  if (warp < kProducerWarps) {
    // ------------------------------------------------------------------ producers
    const int c16 = tid & 7;    // 16-byte chunk inside a 64-channel panel row
    const int psub = tid >> 3;  // this thread covers pairs psub + kRowStep*i of the 64-pair slice
    const int32_t* pa = p.pair_a ? p.pair_a + int64_t(k) * p.pitch : nullptr;
    const int32_t* pb = p.pair_b ? p.pair_b + int64_t(k) * p.pitch : nullptr;
    // Pair indices of kIdxGroup slices are fetched as one batch of independent loads while the previous
    // group's slices are being issued: one memory round trip per group, off the critical path.
    auto load_group = [&](int it0, int (&ia)[kIdxGroup][kPairsPerThread], int (&ib)[kIdxGroup][kPairsPerThread]) {
#pragma unroll
      for (int u = 0; u < kIdxGroup; ++u) {
#pragma unroll
        for (int i = 0; i < kPairsPerThread; ++i) {
          const int64_t q = begin + int64_t(it0 + u) * kSliceK + psub + kRowStep * i;
          int va = -1, vb = -1;
          if (q < end) {
            va = pa ? __ldg(pa + q) : int(q);
            vb = pb ? __ldg(pb + q) : int(q);
          }
          ia[u][i] = va; ib[u][i] = vb;
        }
      }
    };
    // loop invariants: swizzled smem offset of this thread's chunk (pairs psub + 16 i are 2048 B
    // apart, panels 8192 B apart), which panels' chunks lie inside the padded rows
    const uint32_t off0 = sw128_offset(uint32_t(psub), uint32_t(c16));
    const size_t a_row_bytes = size_t(p.ca_pad) * 2, b_row_bytes = size_t(p.cb_pad) * 2;
    const char* a_c = reinterpret_cast<const char*>(p.a) + size_t(a_c0 + c16 * 8) * 2;
    const char* b_c = reinterpret_cast<const char*>(p.b) + size_t(b_c0 + c16 * 8) * 2;
    uint32_t a_ok[2 * MT], b_ok[4];
#pragma unroll
    for (int pn = 0; pn < 2 * MT; ++pn) a_ok[pn] = (a_c0 + pn * 64 + c16 * 8 < p.ca_pad) ? 16u : 0u;
#pragma unroll
    for (int pn = 0; pn < 4; ++pn) b_ok[pn] = (pn < b_panels && b_c0 + pn * 64 + c16 * 8 < p.cb_pad) ? 16u : 0u;
    int ia[kIdxGroup][kPairsPerThread], ib[kIdxGroup][kPairsPerThread], na[kIdxGroup][kPairsPerThread], nb[kIdxGroup][kPairsPerThread];
    if (iters > 0) load_group(0, ia, ib);
    int s = 0;
    uint32_t ph = 0;
    for (int it0 = 0; it0 < iters; it0 += kIdxGroup) {
      if (it0 + kIdxGroup < iters) load_group(it0 + kIdxGroup, na, nb);
#pragma unroll
      for (int u = 0; u < kIdxGroup; ++u) {
        const int it = it0 + u;
        if (it < iters) {
          if (it >= p.stages) mbar_wait(&bars.free_[s], ph ^ 1u);
          const uint32_t sa = smem0 + uint32_t(s) * stage_bytes + off0;
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int i = 0; i < kPairsPerThread; ++i) {
            int va = ia[u][i], vb = ib[u][i];
            const bool ok = va >= 0 && vb >= 0 && va < p.n_a && vb < p.n_b;
            const uint32_t live = ok ? 16u : 0u;
            const char* ga = a_c + (ok ? size_t(va) * a_row_bytes : size_t(0));
            const char* gb = b_c + (ok ? size_t(vb) * b_row_bytes : size_t(0));
            // A: [64 pairs][MT x 128 channels of a] as 64-channel swizzled panels
#pragma unroll
            for (int pn = 0; pn < 2 * MT; ++pn)
              if (pn < 2 * mt_live) cp_async16(sa + pn * 8192 + i * (kRowStep * 128), ga + pn * 128, live & a_ok[pn]);
            // B: [64 pairs][n_tile channels of b]
#pragma unroll
            for (int pn = 0; pn < 4; ++pn)
              if (pn < b_panels) cp_async16(sb + pn * 8192 + i * (kRowStep * 128), gb + pn * 128, live & b_ok[pn]);
          }
          cp_async_arrive_noinc(&bars.full[s]);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
#pragma unroll
      for (int u = 0; u < kIdxGroup; ++u)
#pragma unroll
        for (int i = 0; i < kPairsPerThread; ++i) { ia[u][i] = na[u][i]; ib[u][i] = nb[u][i]; }
    }
    // ------------------------------------------------------------------ epilogue
    // TMEM -> registers (thread = a-channel) -> this warp's smem tile -> global, a warp adding / storing
    // 32 consecutive b-channels of one a-channel row at a time (coalesced stores / reductions)
    if (iters > 0) {
      mbar_wait(&bars.done, 0);
      tc_fence_after();
    }
    float* tile = reinterpret_cast<float*>(smem) + (warp % kEpiWarps) * (32 * kEpiPitch);
    for (int mt = 0; mt < (warp < kEpiWarps ? mt_live : 0); ++mt) {
      const int ca0 = a_c0 + mt * kTileM + (warp & 3) * 32;
      float* out = p.dw + (int64_t(k) * p.c_a + ca0) * p.c_b;
      int rmax = p.c_a - ca0;
      if (rmax > 32) rmax = 32;
      for (int col = (warp >> 2) * 32; col < p.n_tile; col += 64) {  // warps w, w+4: alternate chunks of lane quarter w
        const int ncols = p.n_tile - col < 32 ? p.n_tile - col : 32;  // 16 or 32
        uint32_t acc[32];
        if (iters > 0) {
          const uint32_t taddr = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t(mt * p.acc_stride + col);
          tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
          if (ncols > 16) tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (4 * q < ncols)
            *reinterpret_cast<uint4*>(tile + lane * kEpiPitch + 4 * q) = make_uint4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        __syncwarp();
        const int cb = b_c0 + col + lane;
        if (lane < ncols && cb < p.c_b) {
          float* o = out + cb;
          if (p.use_atomic) {
            for (int r = 0; r < rmax; ++r) atomicAdd(o + int64_t(r) * p.c_b, tile[r * kEpiPitch + lane]);
          } else {
            for (int r = 0; r < rmax; ++r) o[int64_t(r) * p.c_b] = tile[r * kEpiPitch + lane];
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 1, 1);
    // MN-major: 64-channel panels 8192 B apart (LBO), 8-pair groups 1024 B apart (SBO); the K=16 steps of a slice
    // are 2048 B apart = +128 in the descriptor's start-address field
    const uint64_t desc_hi = make_desc_sw128(0, 8192, 1024);
    const uint32_t full0 = smem_u32(&bars.full[0]), free0 = smem_u32(&bars.free_[0]);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait_u32(full0 + 8u * s, ph);
      fence_proxy_async_smem();
      tc_fence_after();
      const uint32_t a_addr = smem0 + uint32_t(s) * stage_bytes, b_addr = a_addr + a_bytes;
      if (elect_one()) {
        const uint64_t bdesc0 = desc_hi | uint64_t((b_addr >> 4) & 0x3fffu);
        for (int mt = 0; mt < mt_live; ++mt) {
          const uint64_t adesc0 = desc_hi | uint64_t(((a_addr + mt * 16384) >> 4) & 0x3fffu);
#pragma unroll
          for (int kk = 0; kk < kSliceK / 16; ++kk)
            mma_bf16(tmem + uint32_t(mt * p.acc_stride), adesc0 + 128 * kk, bdesc0 + 128 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        mma_commit_u32(free0 + 8u * s);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    if (iters > 0 && elect_one()) mma_commit(&bars.done);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, tmem_cols);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int g_force_rblk = 0;  // tuning knob (wfsp_set_option "apply_row_blocks"): 0 = cost model
int g_auto_ksplit = 4;  // wfsp_set_option "apply_k_split": 0 = never split the reduction over a cluster, else the largest split
// wfsp_set_option "apply_bn_fuse": 1 = small launches finish the BatchNorm behind them inside the launch (grid barrier).
// Measured on B200 and left OFF: at 64 events the fused launches take 22 / 29 us against 11.5 / 18.7 us for convolution
// + stand-alone BatchNorm -- the few warps that own blocks fold the statistics serially, while the stand-alone kernel
// spreads the same work over every SM (profiles/r2_experiments.md).
int g_bn_fuse = 0;
int g_split_wide = 0;    // wfsp_set_option "apply_split_wide": split launches take the widest column tiles (fewest MMA instructions)
int g_prep_ctas = 0;     // wfsp_set_option "prep_ctas": CTAs of one weight-preparation launch, 0 = SM count - 52
int g_split_stages = 4;  // wfsp_set_option "apply_split_stages": ring depth of split launches (each CTA walks few slices)

// column tiling of the destination channels: as few tiles as possible (<= 256 columns each) when
// there are enough row tiles to fill the machine, narrower tiles (down to 32 columns) when there are
// not -- a small problem is bound by the serial k-loop of its few CTAs, so spreading the columns over
// idle SMs shortens it; the price (the A tile is gathered once per column tile) is negligible there.
void choose_column_tiles(int n_pad, int64_t live_row_tiles, int& n_tile, int& n_tiles) {
  const int t_min = (n_pad + 255) / 256;
  const int t_max = (n_pad + 31) / 32;
  int64_t want = ceil_div<int64_t>(sm_count(), live_row_tiles > 0 ? live_row_tiles : 1);
  int t = int(want < t_min ? t_min : (want > t_max ? t_max : want));
  n_tile = round_up((n_pad + t - 1) / t, 16);
  n_tiles = (n_pad + n_tile - 1) / n_tile;
}

struct ApplyPlan { int n_pad, kc_pad, c_pad; size_t off_act, total; };
ApplyPlan apply_plan(int kvol, int64_t n_src, int c_red, int c_dst) {
  ApplyPlan a;
  a.n_pad = round_up(c_dst, 16);
  a.kc_pad = round_up(c_red, kSliceK);
  a.c_pad = round_up(c_red, 8);
  a.off_act = align_up(size_t(kvol) * a.n_pad * a.kc_pad * sizeof(__nv_bfloat16), 256);
  a.total = a.off_act + align_up(size_t(n_src) * a.c_pad * sizeof(__nv_bfloat16), 256);
  return a;
}

// stages that fit: prefer two resident CTAs per SM when a 4-deep ring fits in ~110 KB
// `alone`: the launch has at most one CTA per SM anyway (a small, latency-bound problem): take the deepest
// ring that fits, every extra slice in flight shortens the serial walk over (offset, channel slice).
int pick_stages(int stage_bytes, int extra_bytes, bool alone = false) {
  if (!alone && 4 * stage_bytes + extra_bytes <= 110 * 1024) return 4;
  int s = (kSmemMax - extra_bytes) / stage_bytes;
  const int cap = alone ? kMaxStages : 6;
  return s < 2 ? 2 : (s > cap ? cap : s);
}

}  // namespace

void set_force_rblk(int v) { g_force_rblk = v; }
void set_auto_ksplit(int v) { g_auto_ksplit = v == 1 ? 8 : v; }
void set_split_stages(int v) { g_split_stages = v; }
void set_split_wide(int v) { g_split_wide = v; }
void set_prep_ctas(int v) { g_prep_ctas = v; }
void set_bn_fuse(int v) { g_bn_fuse = v; }
unsigned long long* g_trace = nullptr;
void set_trace(unsigned long long* p) { g_trace = p; }

size_t conv_apply_umma_workspace(int kvol, int64_t n_src, int c_red, int c_dst, int split3) {
  return apply_plan(kvol, n_src, split3 ? 3 * c_red : c_red, c_dst).total;
}

int conv_apply_umma_launch(const __nv_bfloat16* act, int64_t n_src, int c_red, const __nv_bfloat16* wt,
                           const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                           const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                           const wfsp_conv_epilogue* ep, cudaStream_t st);

// fp32 in / fp32 out entry: weight preparation + activation cast (one launch), then the tcgen05 kernel.  split3
// (WFSP_MATH_BF16X3): the reduction runs over three bf16 segments per channel -- [hi(a) | hi(a) | lo(a)] against
// [hi(w); lo(w); hi(w)] -- i.e. a w ~ hi hi + hi lo + lo hi accumulated in fp32: fp32-grade results (error ~2^-16 per
// product) on the tensor cores, through the very same kernel with a three times longer reduction.
int conv_apply_umma(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst, void* ws,
                    size_t ws_bytes, const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                    int split3, cudaStream_t st) {
  if (n_dst == 0) return WFSP_OK;
  const int c_eff = split3 ? 3 * c_red : c_red;
  ApplyPlan a = apply_plan(kvol, n_src, c_eff, c_dst);
  if (ws == nullptr || ws_bytes < a.total) return set_error(WFSP_EWORKSPACE, "conv_apply workspace %zu < %zu", ws_bytes, a.total);
  __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + a.off_act);
  {
    const int64_t total = int64_t(kvol) * a.n_pad * a.kc_pad;
    int64_t prep_blocks = ceil_div<int64_t>(total, 256);
    if (prep_blocks > int64_t(sm_count()) * 8) prep_blocks = int64_t(sm_count()) * 8;
    CastJob job{src, act, n_src, c_red, a.c_pad, n_src_dev, split3 ? 1 : 0};
    int64_t cast_blocks = ceil_div<int64_t>(n_src * (a.c_pad >> 3) > 0 ? n_src * (a.c_pad >> 3) : 1, 256);
    if (cast_blocks > int64_t(sm_count()) * 16) cast_blocks = int64_t(sm_count()) * 16;
    prep_and_cast_kernel<<<unsigned(prep_blocks + cast_blocks), 256, 0, st>>>(
        weight, kvol, c_eff, c_dst, transpose_w, wt, a.n_pad, a.kc_pad, int(prep_blocks), job, split3);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }
  return conv_apply_umma_launch(act, n_src, c_eff, wt, bias, nbr, kvol, dst, n_dst, c_dst, n_src_dev, n_dst_dev,
                                n_dst_hint, nullptr, st);
}

// bf16 activations [n_src, round_up(c_red, 8)] and prepared weights in, fp32 [n_dst, c_dst] out
int conv_apply_umma_launch(const __nv_bfloat16* act, int64_t n_src, int c_red, const __nv_bfloat16* wt,
                           const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                           const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                           const wfsp_conv_epilogue* ep, cudaStream_t st) {
  if (n_dst == 0) return WFSP_OK;
  ApplyPlan a = apply_plan(kvol, n_src, c_red, c_dst);
  const int64_t live = (n_dst_hint > 0 && n_dst_hint < n_dst) ? n_dst_hint : n_dst;
  const int64_t row_blocks = ceil_div<int64_t>(live, kTileM);
  const int iters_est = kvol * (a.kc_pad / kSliceK);
  // BatchNorm inside the launch: small launches only (all CTAs co-resident, checked below); otherwise, or when the
  // check fails, the convolution runs as usual and the stand-alone BatchNorm launch follows it right here
  const wfsp_bn_fuse* bn = (ep != nullptr && ep->bn != nullptr && ep->bn_partials != nullptr) ? ep->bn : nullptr;
  bool fuse_bn = bn != nullptr && g_bn_fuse && live <= 8192 && c_dst <= 512 && bn->barrier != nullptr;
  auto unfused = [&]() -> int {
    wfsp_conv_epilogue e2 = *ep;
    e2.bn = nullptr;
    if (int rc = conv_apply_umma_launch(act, n_src, c_red, wt, bias, nbr, kvol, dst, n_dst, c_dst, n_src_dev, n_dst_dev,
                                        n_dst_hint, &e2, st))
      return rc;
    return wfsp_bn_relu_fwd_stats_ex(dst, n_dst, n_dst_dev, n_dst_hint, c_dst, ep->bn_partials, bn->gamma, bn->beta,
                                     bn->running_mean, bn->running_var, bn->momentum, bn->eps, bn->relu, bn->y, bn->y_bf16,
                                     bn->save_mean, bn->save_invstd, bn->dropout, reinterpret_cast<wfsp_stream_t>(st));
  };
  if (bn != nullptr && !fuse_bn) return unfused();
  ApplyParams p{};
  p.src = act; p.n_src = n_src; p.c_pad = a.c_pad; p.wt = wt; p.n_pad = a.n_pad; p.kc_pad = a.kc_pad;
  p.bias = bias; p.nbr = nbr; p.kvol = kvol; p.dst = dst; p.n_dst = n_dst; p.c_dst = c_dst;
  p.n_src_dev = n_src_dev; p.n_dst_dev = n_dst_dev;
  p.trace = g_trace;
  if (ep) {
    p.stats = ep->bn_partials;
    if (ep->bwd_partials) {
      if (!ep->bwd_x || !ep->bwd_mean || !ep->bwd_invstd)
        return set_error(WFSP_EINVAL, "bwd_partials needs bwd_x, bwd_mean and bwd_invstd");
      p.bwd_x = ep->bwd_x; p.bwd_mean = ep->bwd_mean; p.bwd_invstd = ep->bwd_invstd; p.bwd_gamma = ep->bwd_gamma;
      p.bwd_beta = ep->bwd_beta; p.bwd_part = ep->bwd_partials; p.bwd_relu = ep->bwd_relu;
    }
  }
  // Split reduction over a thread-block cluster (small launches only): the largest split that leaves every CTA
  // at least three pipeline slices while the launch still fits one round of the machine.
  int ksplit = 1;
  const int want_split = ep ? ep->k_split : 0;
  if (want_split == 0) {
    const int t_min = (a.n_pad + 255) / 256;
    for (int ks = 8; ks > 1; ks >>= 1)
      if (ks <= g_auto_ksplit && iters_est >= 3 * ks && row_blocks * t_min * ks <= sm_count()) { ksplit = ks; break; }
  } else if (want_split == 2 || want_split == 4 || want_split == 8) {
    ksplit = want_split;
  }
  const size_t block_bytes = size_t(32) * kEpiPitch * sizeof(float);
  int n_tile = 0, n_tiles = 0, nbr_bytes = 0, stage_bytes = 0;
  size_t park_bytes = 0;
  for (;; ksplit = 1) {  // second trip: the split does not fit the shared memory, plain launch
    choose_column_tiles(a.n_pad, row_blocks * ksplit, n_tile, n_tiles);
    if (ksplit > 1 && g_split_wide) {
      // issuing a tcgen05.mma costs about the same whatever its N: a split launch takes the fewest, widest column
      // tiles (its parallelism comes from the split), which also gathers every row once instead of once per tile
      n_tiles = (a.n_pad + 255) / 256;
      n_tile = round_up((a.n_pad + n_tiles - 1) / n_tiles, 16);
    }
    p.n_tile = n_tile;
    p.acc_stride = round_up(n_tile, 32);
    // Row blocking: rblk 128-row blocks per CTA share every weight slice.  Modelled cost of a launch =
    // rounds of CTAs over the SMs x (operand KB a CTA moves + a per-CTA prologue / epilogue term); the weight
    // slice (n_tile * 128 B per stage) is amortised over rblk blocks, the rounds quantise.  Only worth it
    // when the row blocks exceed one round -- small problems keep rblk = 1 and as many CTAs as possible.
    int best_r = 1;
    double best_cost = 1e300;
    const bool tile_loop_kernel = ksplit > 1 || p.bwd_part != nullptr || fuse_bn;  // conv_apply_split_kernel: one row block per CTA
    for (int r = 1; r <= (tile_loop_kernel ? 1 : kMaxRowBlocks); ++r) {
      if (r * p.acc_stride > 512) break;
      const int stage_b = r * kABytes + n_tile * 128;
      const int nbr_b = (nbr != nullptr && kvol <= kNbrStageK) ? r * kTileM * kvol * 4 : 0;
      if (2 * stage_b + nbr_b + 1024 > kSmemMax) break;
      const int64_t ctas = ceil_div<int64_t>(row_blocks, r) * n_tiles;
      const double rounds = double(ceil_div<int64_t>(ctas, sm_count()));
      const double cost = rounds * (double(iters_est) * (r * 16.0 + n_tile / 8.0) + 48.0 * r + 48.0);
      if (cost < best_cost * 0.97) { best_cost = cost; best_r = r; }
    }
    if (!tile_loop_kernel && g_force_rblk > 0 && g_force_rblk <= kMaxRowBlocks && g_force_rblk * p.acc_stride <= 512) best_r = g_force_rblk;
    p.rblk = best_r;
    stage_bytes = p.rblk * kABytes + n_tile * 128;
    p.staged = (nbr != nullptr && kvol <= kNbrStageK) ? 1 : 0;
    nbr_bytes = p.staged ? p.rblk * kTileM * kvol * 4 : 0;
    const bool alone = ceil_div<int64_t>(row_blocks, p.rblk) * n_tiles * ksplit <= sm_count();
    park_bytes = 0;
    if (ksplit == 1) {
      p.stages = pick_stages(stage_bytes, nbr_bytes + 1024, alone);
      break;
    }
    // a split CTA walks only a few slices: a shallow ring; behind it the slots its peers push their partial blocks
    // into (ks slots for every block this CTA finishes); the ring must also hold one block per epilogue warp
    const int units = p.rblk * 4 * ((n_tile + 31) / 32);
    park_bytes = size_t(ceil_div(units, ksplit)) * ksplit * block_bytes;
    int stages = g_split_stages < 2 ? 2 : g_split_stages;
    while (stages > 2 && size_t(stages) * stage_bytes + nbr_bytes + park_bytes + 1024 > size_t(kSmemMax)) --stages;
    while (size_t(stages) * stage_bytes < kEpiWarps * block_bytes) ++stages;
    p.stages = stages;
    if (size_t(stages) * stage_bytes + nbr_bytes + park_bytes + 1024 <= size_t(kSmemMax)) break;
  }
  p.ksplit = ksplit;
  p.nbr_bytes = nbr_bytes;
  const size_t smem = size_t(p.stages) * stage_bytes + nbr_bytes + park_bytes + 1024;
  if (smem > size_t(227) * 1024) return set_error(WFSP_EUNSUPPORTED, "conv_apply tile needs %zu B of shared memory", smem);
  // grid: the tiles of the EXPECTED live rows plus half as many again (the kernel loops over whatever lies beyond)
  const int64_t tile_rows = int64_t(kTileM) * p.rblk;
  int64_t grid_x = ceil_div<int64_t>(n_dst, tile_rows);
  const bool tile_loop_kernel = ksplit > 1 || p.bwd_part != nullptr || fuse_bn;
  if (tile_loop_kernel && live < n_dst) {
    // (a fused BatchNorm needs every CTA resident at once: one spare tile instead of half as many again)
    const int64_t want = fuse_bn ? ceil_div<int64_t>(live, tile_rows) + 1 : ceil_div<int64_t>(live + live / 2, tile_rows);
    if (want < grid_x) grid_x = want;
  }
  const dim3 grid{unsigned(grid_x), unsigned(n_tiles), unsigned(ksplit)};
  if (fuse_bn) {
    // co-residency of the whole grid (the barrier spins): clusters for the split kernel, CTAs otherwise
    int fit = 0;
    if (ksplit > 1) {
      WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_split_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
      cudaLaunchConfig_t qc = {};
      qc.gridDim = grid; qc.blockDim = dim3(kApplyThreads, 1, 1); qc.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 1; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = unsigned(ksplit);
      qc.attrs = qa; qc.numAttrs = 1;
      int clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&clusters, conv_apply_split_kernel<1, true>, &qc) != cudaSuccess) { cudaGetLastError(); clusters = 0; }
      fit = int64_t(clusters) >= grid_x * n_tiles ? 1 : 0;
    } else {
      WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_split_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_apply_split_kernel<1, false>, kApplyThreads, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
      fit = int64_t(per_sm) * sm_count() >= grid_x * n_tiles ? 1 : 0;
    }
    if (!fit) return unfused();
    p.bn_fuse = 1; p.bn_relu = bn->relu; p.bn_gamma = bn->gamma; p.bn_beta = bn->beta;
    p.bn_rmean = bn->running_mean; p.bn_rvar = bn->running_var; p.bn_save_mean = bn->save_mean; p.bn_save_invstd = bn->save_invstd;
    p.bn_eps = bn->eps; p.bn_momentum = bn->momentum; p.bn_y32 = bn->y; p.bn_y16 = static_cast<__nv_bfloat16*>(bn->y_bf16);
    p.bn_bar = static_cast<unsigned*>(bn->barrier);
    p.bn_drop = make_drop(bn->dropout);
  }
  if (tile_loop_kernel && ksplit == 1) {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_split_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    conv_apply_split_kernel<1, false><<<grid, kApplyThreads, smem, st>>>(p);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  if (ksplit > 1) {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_split_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kApplyThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = unsigned(ksplit);
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WFSP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_apply_split_kernel<1, true>, p));
    count_launches(1);
    return WFSP_OK;
  }
  switch (p.rblk) {
#define WFSP_LAUNCH_APPLY(RB)                                                                                        \
    case RB:                                                                                                         \
      WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_umma_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           kSmemOptIn));                                                             \
      conv_apply_umma_kernel<RB><<<grid, kApplyThreads, smem, st>>>(p);                                            \
      break;
    WFSP_LAUNCH_APPLY(1) WFSP_LAUNCH_APPLY(2) WFSP_LAUNCH_APPLY(3) WFSP_LAUNCH_APPLY(4)
#undef WFSP_LAUNCH_APPLY
    default: return set_error(WFSP_EINVAL, "bad row blocking %d", p.rblk);
  }
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

size_t conv_wgrad_umma_workspace(int, int64_t n_a, int c_a, int64_t n_b, int c_b, int64_t, int split3) {
  const size_t one = align_up(size_t(n_a) * round_up(c_a, 8) * 2, 256) + align_up(size_t(n_b) * round_up(c_b, 8) * 2, 256);
  return split3 ? 2 * one : one;
}

int conv_wgrad_umma_launch(const __nv_bfloat16* a16, int64_t n_a, int c_a, const __nv_bfloat16* b16, int64_t n_b, int c_b,
                           const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol,
                           int64_t pitch, float* d_weight, int accumulate, const int32_t* n_a_dev, int64_t pairs_hint,
                           cudaStream_t st);

// split3: d_weight = hi(a)^T hi(b) + hi(a)^T lo(b) + lo(a)^T hi(b), three accumulating launches of the same kernel
int conv_wgrad_umma(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, void* ws, size_t ws_bytes, const int32_t* n_a_dev,
                    const int32_t* n_b_dev, int64_t pairs_hint, int split3, cudaStream_t st) {
  const size_t need = conv_wgrad_umma_workspace(kvol, n_a, c_a, n_b, c_b, pitch, split3);
  if (need > 0 && (ws == nullptr || ws_bytes < need))
    return set_error(WFSP_EWORKSPACE, "conv_wgrad workspace %zu < %zu", ws_bytes, need);
  const int ca_pad = round_up(c_a, 8), cb_pad = round_up(c_b, 8);
  const size_t a_bytes = align_up(size_t(n_a) * ca_pad * 2, 256), b_bytes = align_up(size_t(n_b) * cb_pad * 2, 256);
  char* w8 = static_cast<char*>(ws);
  __nv_bfloat16* a16 = reinterpret_cast<__nv_bfloat16*>(w8);
  __nv_bfloat16* b16 = reinterpret_cast<__nv_bfloat16*>(w8 + a_bytes);
  CastJob ja{a, a16, n_a, c_a, ca_pad, n_a_dev, split3 ? 2 : 0}, jb{b, b16, n_b, c_b, cb_pad, n_b_dev, split3 ? 2 : 0};
  if (int rc = launch_cast(ja, &jb, st)) return rc;
  if (!split3)
    return conv_wgrad_umma_launch(a16, n_a, c_a, b16, n_b, c_b, pair_a, pair_b, pair_num, kvol, pitch, d_weight, accumulate,
                                  n_a_dev, pairs_hint, st);
  __nv_bfloat16* a_lo = reinterpret_cast<__nv_bfloat16*>(w8 + a_bytes + b_bytes);
  __nv_bfloat16* b_lo = reinterpret_cast<__nv_bfloat16*>(w8 + 2 * a_bytes + b_bytes);
  CastJob la{a, a_lo, n_a, c_a, ca_pad, n_a_dev, 3}, lb{b, b_lo, n_b, c_b, cb_pad, n_b_dev, 3};
  if (int rc = launch_cast(la, &lb, st)) return rc;
  if (int rc = conv_wgrad_umma_launch(a16, n_a, c_a, b16, n_b, c_b, pair_a, pair_b, pair_num, kvol, pitch, d_weight, accumulate,
                                      n_a_dev, pairs_hint, st)) return rc;
  if (int rc = conv_wgrad_umma_launch(a16, n_a, c_a, b_lo, n_b, c_b, pair_a, pair_b, pair_num, kvol, pitch, d_weight, 1,
                                      n_a_dev, pairs_hint, st)) return rc;
  return conv_wgrad_umma_launch(a_lo, n_a, c_a, b16, n_b, c_b, pair_a, pair_b, pair_num, kvol, pitch, d_weight, 1,
                                n_a_dev, pairs_hint, st);
}

// bf16 rows in (pitches round_up(c, 8)), fp32 d_weight out
int conv_wgrad_umma_launch(const __nv_bfloat16* a16, int64_t n_a, int c_a, const __nv_bfloat16* b16, int64_t n_b, int c_b,
                           const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol,
                           int64_t pitch, float* d_weight, int accumulate, const int32_t* n_a_dev, int64_t pairs_hint,
                           cudaStream_t st) {
  if (n_a == 0 || n_b == 0 || (pair_a != nullptr && pitch == 0)) {  // no pairs: the gradient is zero
    if (!accumulate) WFSP_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, size_t(kvol) * c_a * c_b * sizeof(float), st));
    return WFSP_OK;
  }
  WgradParams p{};
  p.ca_pad = round_up(c_a, 8);
  p.cb_pad = round_up(c_b, 8);
  p.n_a_dev = n_a_dev;
  p.a = a16; p.n_a = n_a; p.c_a = c_a; p.b = b16; p.n_b = n_b; p.c_b = c_b;
  p.pair_a = pair_a; p.pair_b = pair_b; p.pair_num = pair_num; p.kvol = kvol; p.pitch = pitch; p.dw = d_weight;
  // pairs per offset that bound the split of the reduction: the caller's hint (graph path, where only
  // capacities are known on the host) or the capacity itself
  int64_t rows = pair_a ? pitch : n_a;
  if (pairs_hint > 0 && pairs_hint < rows) rows = pairs_hint;
  // A short pair list (a few pipeline slices) makes the launch latency bound: its cost is the epilogue of
  // the few CTAs, so the d_weight tile is cut small (128 x 64) and spread over many CTAs.  Long lists take
  // wide tiles (up to 256 x 256) that reuse every gathered row as much as TMEM allows.
  const bool few_pairs = rows <= 32 * kSliceK;
  int n_tiles = (c_b + 255) / 256;
  if (few_pairs) n_tiles = (c_b + 63) / 64;
  p.n_tile = round_up((c_b + n_tiles - 1) / n_tiles, 16);
  p.acc_stride = round_up(p.n_tile, 32);
  // two 128-channel blocks of `a` per CTA when `a` has more than 128 channels and both accumulators fit
  const int mt = (!few_pairs && c_a > kTileM && 2 * p.acc_stride <= 512) ? 2 : 1;
  p.m_groups = ceil_div(c_a, kTileM * mt);
  const int tiles = p.m_groups * n_tiles;
  const int b_panels = (p.n_tile + 63) / 64;
  const int stage_bytes = 2 * mt * 8192 + b_panels * 8192;
  p.stages = pick_stages(stage_bytes, 1024);
  const size_t smem = size_t(p.stages) * stage_bytes + 1024;
  // Split the pair list so that the CTAs fill whole rounds of the machine.  Modelled cost of a launch =
  // rounds x (pipeline slices of one CTA + a fixed term for its prologue and the atomic epilogue).
  int nsplit = 1;
  if (rows > 0) {
    const int64_t slots = int64_t(sm_count()) * ((2 * (smem + 2048) <= size_t(227) * 1024 && 2 * mt * p.acc_stride <= 512) ? 2 : 1);
    const int64_t units = int64_t(tiles) * kvol;
    int64_t maxs = ceil_div<int64_t>(rows, 256);
    if (maxs > 64) maxs = 64;
    if (maxs * kvol > 65535) maxs = 65535 / kvol;
    double best = 1e300;
    for (int64_t ns = 1; ns <= maxs; ++ns) {
      const double rounds = double(ceil_div<int64_t>(units * ns, slots));
      const double cost = rounds * (double(ceil_div<int64_t>(rows, ns * kSliceK)) + 24.0);
      if (cost < best * 0.98) { best = cost; nsplit = int(ns); }
    }
  }
  p.nsplit = nsplit;
  p.chunk = round_up(int(ceil_div<int64_t>(rows > 0 ? rows : 1, nsplit)), kSliceK);
  p.use_atomic = (nsplit > 1 || accumulate) ? 1 : 0;
  if (p.use_atomic && !accumulate)
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, size_t(kvol) * c_a * c_b * sizeof(float), st));
  dim3 grid(unsigned(tiles), unsigned(kvol * nsplit));
  if (mt == 2) {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    conv_wgrad_umma_kernel<2><<<grid, kThreads, smem, st>>>(p);
  } else {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptIn));
    conv_wgrad_umma_kernel<1><<<grid, kThreads, smem, st>>>(p);
  }
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

size_t prepared_weight_bytes(int kvol, int c_red, int c_dst) {
  return align_up(size_t(kvol) * round_up(c_dst, 16) * round_up(c_red, kSliceK) * sizeof(__nv_bfloat16), 256);
}

int prep_weights_batch(const wfsp_prep_job* jobs, int n_jobs, cudaStream_t st) {
  for (int j0 = 0; j0 < n_jobs; j0 += kMaxPrepJobs) {
    PrepBatch b{};
    b.n = n_jobs - j0 < kMaxPrepJobs ? n_jobs - j0 : kMaxPrepJobs;
    int64_t chunks[kMaxPrepJobs], all = 0;
    for (int j = 0; j < b.n; ++j) {
      const wfsp_prep_job& q = jobs[j0 + j];
      if (q.weight == nullptr || q.out == nullptr || q.kvol < 1 || q.c_red < 1 || q.c_dst < 1)
        return set_error(WFSP_EINVAL, "bad weight-preparation job %d", j0 + j);
      if ((reinterpret_cast<uintptr_t>(q.out) & 15) != 0)
        return set_error(WFSP_EINVAL, "prepared-weight buffer of job %d is not 16-byte aligned", j0 + j);
      chunks[j] = int64_t(q.kvol) * round_up(q.c_dst, 16) * (round_up(q.c_red, kSliceK) >> 3);
      all += chunks[j];
    }
    // about one CTA per SM in total, shared out by size (at least one each, never more than the job has chunks for)
    // (default: about a third of the SMs stay free -- a convolution CTA takes a whole SM's shared memory, so it
    // cannot share one with even a single preparation CTA; 64-event step: first convolution at 14.5 instead of 18.7 us)
    const int64_t budget = g_prep_ctas > 0 ? g_prep_ctas : (sm_count() > 104 ? sm_count() - 52 : sm_count() / 2 + 1);
    int blocks = 0;
    for (int j = 0; j < b.n; ++j) {
      const wfsp_prep_job& q = jobs[j0 + j];
      int64_t nb = (budget * chunks[j] + all / 2) / all;
      const int64_t most = ceil_div<int64_t>(chunks[j], 256);
      if (nb > most) nb = most;
      if (nb < 1) nb = 1;
      b.job[j] = PrepJob{q.weight, static_cast<__nv_bfloat16*>(q.out), q.kvol, q.c_red, q.c_dst, q.transpose_w, blocks, int(nb)};
      blocks += int(nb);
    }
    prep_weights_batch_kernel<<<unsigned(blocks), 256, 0, st>>>(b);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }
  return WFSP_OK;
}

int cast_rows_bf16(const float* src, int64_t n, const int32_t* n_dev, int c, void* dst16, cudaStream_t st) {
  CastJob j{src, static_cast<__nv_bfloat16*>(dst16), n, c, round_up(c, 8), n_dev, 0};
  return launch_cast(j, nullptr, st);
}

}  // namespace wfsp
