// tcgen05 (5th-gen tensor core) gather-GEMM kernels, WFSP_MATH_BF16: bf16 operands, fp32
// accumulation in tensor memory.  They replace upstream indiceConv / indiceConvBackward
// (reference call sites src/models/SPConvBlocks.py:498-502 etc.).
//
// Both kernels are warp-specialised (160 threads):
//   warps 0-3  producers: issue 16-byte cp.async copies of gathered bf16 rows and of the weight /
//              gradient slice straight into 128-byte-swizzled shared-memory tiles, then
//              cp.async.mbarrier.arrive.noinc on the stage's "full" barrier and move on -- they
//              never wait for data, so up to `stages` slices are in flight per CTA;
//   warp 4     one elected lane waits on "full", issues 4 tcgen05.mma (K=16 each) accumulating in
//              TMEM and tcgen05.commit's to the stage's "free" barrier;
//   warps 0-3  epilogue after the last commit: tcgen05.ld -> registers -> global.
//
// apply kernel (forward, dgrad, inverse forward, inverse dgrad) -- output-stationary implicit GEMM:
//   a CTA owns 128 destination rows and walks (active kernel offset k) x (64-channel slice):
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

constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;
constexpr int kTileM = 128;   // destination rows (apply) / a-channels (wgrad) per CTA
constexpr int kSliceK = 64;   // reduction elements per stage (one 128 B swizzle row of bf16)
constexpr int kABytes = kTileM * 128;
constexpr int kMaxStages = 8;
constexpr int kNbrStageK = 32;  // kernel volumes up to this keep the CTA's neighbour tile in smem
constexpr int kSmemMax = 200 * 1024;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- weight preparation: fp32 [kvol][c_red][c_dst] (or transposed) -> bf16 [kvol][n_pad][kc_pad]
__global__ void __launch_bounds__(256) prep_weight_kernel(const float* __restrict__ w, int kvol, int c_red,
                                                          int c_dst, int transpose_w, __nv_bfloat16* __restrict__ wt,
                                                          int n_pad, int kc_pad) {
  const int64_t total = int64_t(kvol) * n_pad * kc_pad;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int c = int(i % kc_pad);
    int64_t t = i / kc_pad;
    int n = int(t % n_pad);
    int k = int(t / n_pad);
    float v = 0.f;
    if (c < c_red && n < c_dst) {
      const float* wk = w + int64_t(k) * c_red * c_dst;
      v = transpose_w ? wk[int64_t(n) * c_red + c] : wk[int64_t(c) * c_dst + n];
    }
    wt[i] = __float2bfloat16_rn(v);
  }
}

// ---- activation cast: fp32 [n][c] -> bf16 [n][c_pad], zero padded; one 16-byte chunk per thread
struct CastJob { const float* src; __nv_bfloat16* dst; int64_t n; int c, c_pad; const int32_t* n_dev; };

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
    const float* s = j.src + row * j.c + col;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (col + e < j.c) ? __ldg(s + e) : 0.f;
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
                                                            int kc_pad, int prep_blocks, CastJob job) {
  if (int(blockIdx.x) < prep_blocks) {
    const int64_t total = int64_t(kvol) * n_pad * kc_pad;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(prep_blocks) * blockDim.x) {
      int c = int(i % kc_pad);
      int64_t t = i / kc_pad;
      int n = int(t % n_pad);
      int k = int(t / n_pad);
      float v = 0.f;
      if (c < c_red && n < c_dst) {
        const float* wk = w + int64_t(k) * c_red * c_dst;
        v = transpose_w ? wk[int64_t(n) * c_red + c] : wk[int64_t(c) * c_dst + n];
      }
      wt[i] = __float2bfloat16_rn(v);
    }
    return;
  }
  const int64_t n = job.n_dev ? int64_t(*job.n_dev) : job.n;
  const int cpr = job.c_pad >> 3;
  const int64_t chunks = n * cpr;
  const int64_t nb = int64_t(gridDim.x) - prep_blocks;
  for (int64_t i = (int64_t(blockIdx.x) - prep_blocks) * blockDim.x + threadIdx.x; i < chunks; i += nb * blockDim.x) {
    const int64_t row = i / cpr;
    const int col = int(i - row * cpr) << 3;
    const float* s = job.src + row * job.c + col;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (col + e < job.c) ? __ldg(s + e) : 0.f;
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(job.dst + row * job.c_pad + col) = u;
  }
}

int launch_cast(const CastJob& a, const CastJob* b, cudaStream_t st) {
  CastJob j1 = b ? *b : CastJob{nullptr, nullptr, 0, 8, 8, nullptr};
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

__device__ __forceinline__ void init_pipe(PipeBarriers& b) {
  for (int s = 0; s < kMaxStages; ++s) {
    mbar_init(&b.full[s], kProducerThreads);
    mbar_init(&b.free_[s], 1);
  }
  mbar_init(&b.done, 1);
  fence_mbar_init();
}

struct ApplyParams {
  const __nv_bfloat16* src; int64_t n_src; int c_pad;
  const __nv_bfloat16* wt; int n_pad, kc_pad;
  const float* bias; const int32_t* nbr; int kvol;
  float* dst; int64_t n_dst; int c_dst;
  int n_tile, stages;
  const int32_t* n_src_dev; const int32_t* n_dst_dev;
};

__global__ void __launch_bounds__(kThreads) conv_apply_umma_kernel(const ApplyParams pp) {
  ApplyParams p = pp;
  if (p.n_src_dev) p.n_src = *p.n_src_dev;
  if (p.n_dst_dev) p.n_dst = *p.n_dst_dev;
  if (int64_t(blockIdx.x) * kTileM >= p.n_dst) return;  // capacity-sized grid: nothing live in this tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ PipeBarriers bars;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_active[WFSP_MAX_KVOL / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = int64_t(blockIdx.x) * kTileM;
  const int n0 = blockIdx.y * p.n_tile;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kABytes + uint32_t(p.n_tile) * 128u;
  int32_t* s_nbr = reinterpret_cast<int32_t*>(smem + uint32_t(p.stages) * stage_bytes);
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(p.n_tile));
  const bool staged = p.nbr != nullptr && p.kvol <= kNbrStageK;

  for (int i = tid; i < WFSP_MAX_KVOL / 32; i += kThreads) s_active[i] = 0;
  if (tid == 0) init_pipe(bars);
  if (warp == 4) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // neighbour tile of this CTA: which kernel offsets are active, and (small kernels) a smem copy
  if (p.nbr) {
    const int total = kTileM * p.kvol;
    int64_t rows_left = p.n_dst - row0;
    if (rows_left > kTileM) rows_left = kTileM;
    const int limit = int(rows_left) * p.kvol;
    const int32_t* base = p.nbr + row0 * p.kvol;
    for (int i = tid; i < total; i += kThreads) {
      int v = i < limit ? __ldg(base + i) : -1;
      if (v >= p.n_src) v = -1;
      if (staged) s_nbr[i] = v;
      if (v >= 0) {
        const int k = i % p.kvol;
        atomicOr(&s_active[k >> 5], 1u << (k & 31));
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

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int c16 = tid & 7;    // 16-byte chunk inside the 128-byte slice row
    const int rsub = tid >> 3;  // this thread covers tile rows rsub + 16*i
    int it = 0;
    for (int kw = 0; kw < (p.kvol + 31) / 32; ++kw) {
      uint32_t mask = s_active[kw];
      while (mask) {
        const int k = kw * 32 + __ffs(mask) - 1;
        mask &= mask - 1;
        int rows[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rsub + 16 * i;
          int v;
          if (!p.nbr) {
            v = (row0 + r < p.n_dst) ? int(row0 + r) : -1;
          } else if (staged) {
            v = s_nbr[r * p.kvol + k];
          } else {
            v = (row0 + r < p.n_dst) ? __ldg(p.nbr + (row0 + r) * p.kvol + k) : -1;
            if (v >= p.n_src) v = -1;
          }
          rows[i] = v;
        }
        const __nv_bfloat16* wk = p.wt + (int64_t(k) * p.n_pad + n0) * p.kc_pad;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          if (it >= p.stages) mbar_wait(&bars.free_[s], uint32_t((it / p.stages - 1) & 1));
          const uint32_t sa = smem_u32(smem + uint32_t(s) * stage_bytes);
          const uint32_t sb = sa + kABytes;
          const int col = kb * kSliceK + c16 * 8;
          const bool col_ok = col < p.c_pad;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = col_ok && rows[i] >= 0;
            const __nv_bfloat16* g = ok ? p.src + int64_t(rows[i]) * p.c_pad + col : p.src;
            cp_async16(sa + sw128_offset(uint32_t(rsub + 16 * i), uint32_t(c16)), g, ok ? 16u : 0u);
          }
          const __nv_bfloat16* wb = wk + kb * kSliceK + c16 * 8;
          for (int n = rsub; n < p.n_tile; n += 16) {
            const bool ok = n0 + n < p.n_pad;  // the last column tile may overhang the padded weights
            cp_async16(sb + sw128_offset(uint32_t(n), uint32_t(c16)), ok ? wb + int64_t(n) * p.kc_pad : wb,
                       ok ? 16u : 0u);
          }
          cp_async_arrive_noinc(&bars.full[s]);
        }
      }
    }
    // ------------------------------------------------------------------ epilogue
    if (total_iters > 0) {
      mbar_wait(&bars.done, 0);
      tc_fence_after();
    }
    const int64_t r = row0 + warp * 32 + lane;
    float* out = p.dst + r * p.c_dst;
    const bool vec_ok = (p.c_dst % 4) == 0;
    for (int col = 0; col < p.n_tile; col += 16) {
      uint32_t acc[16];
      if (total_iters > 0) {
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + uint32_t(col), acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0u;
      }
      if (r < p.n_dst) {
        const int c = n0 + col;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = c + 4 * q;
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[e] = __uint_as_float(acc[4 * q + e]);
            if (p.bias && cc + e < p.c_dst) f[e] += __ldg(p.bias + cc + e);
          }
          if (vec_ok && cc + 3 < p.c_dst) {
            *reinterpret_cast<float4*>(out + cc) = make_float4(f[0], f[1], f[2], f[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (cc + e < p.c_dst) out[cc + e] = f[e];
          }
        }
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 0, 0);
    for (int it = 0; it < total_iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(&bars.full[s], uint32_t((it / p.stages) & 1));
      fence_proxy_async_smem();
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + uint32_t(s) * stage_bytes), b_addr = a_addr + kABytes;
#pragma unroll
      for (int kk = 0; kk < kSliceK / 16; ++kk) {
        const uint64_t adesc = make_desc_sw128(a_addr + kk * 32, 16, 1024);
        const uint64_t bdesc = make_desc_sw128(b_addr + kk * 32, 16, 1024);
        mma_bf16(tmem, adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
      }
      mma_commit(&bars.free_[s]);
    }
    if (total_iters > 0) mma_commit(&bars.done);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

struct WgradParams {
  const __nv_bfloat16* a; int64_t n_a; int c_a, ca_pad;
  const __nv_bfloat16* b; int64_t n_b; int c_b, cb_pad;
  const int32_t* pair_a; const int32_t* pair_b; const int32_t* pair_num;
  int kvol; int64_t pitch;
  float* dw;
  int n_tile, m_tiles, nsplit, stages, use_atomic;
  int64_t chunk;
  const int32_t* n_a_dev;
};

__global__ void __launch_bounds__(kThreads) conv_wgrad_umma_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ PipeBarriers bars;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.y / p.nsplit, split = blockIdx.y % p.nsplit;
  const int mt = blockIdx.x % p.m_tiles, nt = blockIdx.x / p.m_tiles;
  const int a_c0 = mt * kTileM, b_c0 = nt * p.n_tile;
  const int64_t n_pairs = p.pair_num ? int64_t(p.pair_num[k]) : (p.n_a_dev ? int64_t(*p.n_a_dev) : p.n_a);
  const int64_t begin = int64_t(split) * p.chunk;
  int64_t end = begin + p.chunk;
  if (end > n_pairs) end = n_pairs;
  if (begin >= n_pairs && (split > 0 || p.use_atomic)) return;  // nothing to add (uniform per CTA)
  const int iters = begin < end ? int((end - begin + kSliceK - 1) / kSliceK) : 0;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_panels = (p.n_tile + 63) / 64;
  const uint32_t a_bytes = 2u * 8192u;
  const uint32_t stage_bytes = a_bytes + uint32_t(b_panels) * 8192u;
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(p.n_tile));

  if (tid == 0) init_pipe(bars);
  if (warp == 4) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int c16 = tid & 7;    // 16-byte chunk inside a 64-channel panel row
    const int psub = tid >> 3;  // this thread covers pairs psub + 16*i of the 64-pair slice
    const int32_t* pa = p.pair_a ? p.pair_a + int64_t(k) * p.pitch : nullptr;
    const int32_t* pb = p.pair_b ? p.pair_b + int64_t(k) * p.pitch : nullptr;
    auto load_idx = [&](int it, int (&ia)[4], int (&ib)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t q = begin + int64_t(it) * kSliceK + psub + 16 * i;
        int va = -1, vb = -1;
        if (q < end) {
          va = pa ? __ldg(pa + q) : int(q);
          vb = pb ? __ldg(pb + q) : int(q);
        }
        if (va < 0 || vb < 0 || va >= p.n_a || vb >= p.n_b) { va = -1; vb = -1; }
        ia[i] = va; ib[i] = vb;
      }
    };
    int ia[4], ib[4], na[4], nb[4];
    if (iters > 0) load_idx(0, ia, ib);
    for (int it = 0; it < iters; ++it) {
      if (it + 1 < iters) load_idx(it + 1, na, nb);  // prefetch the next slice's pair indices
      const int s = it % p.stages;
      if (it >= p.stages) mbar_wait(&bars.free_[s], uint32_t((it / p.stages - 1) & 1));
      const uint32_t sa = smem_u32(smem + uint32_t(s) * stage_bytes);
      const uint32_t sb = sa + a_bytes;
      // A: [64 pairs][128 channels of a] as two 64-channel swizzled panels
#pragma unroll
      for (int panel = 0; panel < 2; ++panel) {
        const int col = a_c0 + panel * 64 + c16 * 8;
        const bool col_ok = col < p.ca_pad;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = col_ok && ia[i] >= 0;
          const __nv_bfloat16* g = ok ? p.a + int64_t(ia[i]) * p.ca_pad + col : p.a;
          cp_async16(sa + panel * 8192 + sw128_offset(uint32_t(psub + 16 * i), uint32_t(c16)), g, ok ? 16u : 0u);
        }
      }
      // B: [64 pairs][n_tile channels of b]
      for (int panel = 0; panel < b_panels; ++panel) {
        const int col = b_c0 + panel * 64 + c16 * 8;
        const bool col_ok = col < p.cb_pad;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = col_ok && ib[i] >= 0;
          const __nv_bfloat16* g = ok ? p.b + int64_t(ib[i]) * p.cb_pad + col : p.b;
          cp_async16(sb + panel * 8192 + sw128_offset(uint32_t(psub + 16 * i), uint32_t(c16)), g, ok ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(&bars.full[s]);
#pragma unroll
      for (int i = 0; i < 4; ++i) { ia[i] = na[i]; ib[i] = nb[i]; }
    }
    // ------------------------------------------------------------------ epilogue
    if (iters > 0) {
      mbar_wait(&bars.done, 0);
      tc_fence_after();
    }
    const int ca = a_c0 + warp * 32 + lane;
    float* out = p.dw + (int64_t(k) * p.c_a + ca) * p.c_b;
    for (int col = 0; col < p.n_tile; col += 16) {
      uint32_t acc[16];
      if (iters > 0) {
        tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + uint32_t(col), acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0u;
      }
      if (ca < p.c_a) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int cb = b_c0 + col + e;
          if (cb < p.c_b) {
            if (p.use_atomic) atomicAdd(out + cb, __uint_as_float(acc[e]));
            else out[cb] = __uint_as_float(acc[e]);
          }
        }
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 1, 1);
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(&bars.full[s], uint32_t((it / p.stages) & 1));
      fence_proxy_async_smem();
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + uint32_t(s) * stage_bytes), b_addr = a_addr + a_bytes;
#pragma unroll
      for (int kk = 0; kk < kSliceK / 16; ++kk) {
        // MN-major: 64-channel panels 8192 B apart (LBO), 8-pair groups 1024 B apart (SBO)
        const uint64_t adesc = make_desc_sw128(a_addr + kk * 2048, 8192, 1024);
        const uint64_t bdesc = make_desc_sw128(b_addr + kk * 2048, 8192, 1024);
        mma_bf16(tmem, adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
      }
      mma_commit(&bars.free_[s]);
    }
    if (iters > 0) mma_commit(&bars.done);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

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
int pick_stages(int stage_bytes, int extra_bytes) {
  if (4 * stage_bytes + extra_bytes <= 110 * 1024) return 4;
  int s = (kSmemMax - extra_bytes) / stage_bytes;
  return s < 2 ? 2 : (s > 6 ? 6 : s);
}

}  // namespace

size_t conv_apply_umma_workspace(int kvol, int64_t n_src, int c_red, int c_dst) {
  return apply_plan(kvol, n_src, c_red, c_dst).total;
}

int conv_apply_umma(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst, void* ws,
                    size_t ws_bytes, const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                    cudaStream_t st) {
  if (n_dst == 0) return WFSP_OK;
  ApplyPlan a = apply_plan(kvol, n_src, c_red, c_dst);
  const int64_t live = (n_dst_hint > 0 && n_dst_hint < n_dst) ? n_dst_hint : n_dst;
  int n_tile, n_tiles;
  choose_column_tiles(a.n_pad, ceil_div<int64_t>(live, kTileM), n_tile, n_tiles);
  if (ws == nullptr || ws_bytes < a.total) return set_error(WFSP_EWORKSPACE, "conv_apply workspace %zu < %zu", ws_bytes, a.total);
  __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + a.off_act);
  {
    const int64_t total = int64_t(kvol) * a.n_pad * a.kc_pad;
    int64_t prep_blocks = ceil_div<int64_t>(total, 256);
    if (prep_blocks > int64_t(sm_count()) * 8) prep_blocks = int64_t(sm_count()) * 8;
    CastJob job{src, act, n_src, c_red, a.c_pad, n_src_dev};
    int64_t cast_blocks = ceil_div<int64_t>(n_src * (a.c_pad >> 3) > 0 ? n_src * (a.c_pad >> 3) : 1, 256);
    if (cast_blocks > int64_t(sm_count()) * 16) cast_blocks = int64_t(sm_count()) * 16;
    prep_and_cast_kernel<<<unsigned(prep_blocks + cast_blocks), 256, 0, st>>>(
        weight, kvol, c_red, c_dst, transpose_w, wt, a.n_pad, a.kc_pad, int(prep_blocks), job);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }

  ApplyParams p{act, n_src, a.c_pad, wt, a.n_pad, a.kc_pad, bias, nbr, kvol, dst, n_dst, c_dst, n_tile, 2,
                n_src_dev, n_dst_dev};
  const int stage_bytes = kABytes + n_tile * 128;
  const int nbr_bytes = (nbr != nullptr && kvol <= kNbrStageK) ? kTileM * kvol * 4 : 0;
  p.stages = pick_stages(stage_bytes, nbr_bytes + 1024);
  const size_t smem = size_t(p.stages) * stage_bytes + nbr_bytes + 1024;
  dim3 grid(unsigned(ceil_div<int64_t>(n_dst, kTileM)), unsigned(n_tiles));
  WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_apply_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  conv_apply_umma_kernel<<<grid, kThreads, smem, st>>>(p);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

size_t conv_wgrad_umma_workspace(int, int64_t n_a, int c_a, int64_t n_b, int c_b, int64_t) {
  return align_up(size_t(n_a) * round_up(c_a, 8) * 2, 256) + align_up(size_t(n_b) * round_up(c_b, 8) * 2, 256);
}

int conv_wgrad_umma(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, void* ws, size_t ws_bytes, const int32_t* n_a_dev,
                    const int32_t* n_b_dev, int64_t pairs_hint, cudaStream_t st) {
  const size_t need = conv_wgrad_umma_workspace(kvol, n_a, c_a, n_b, c_b, pitch);
  if (need > 0 && (ws == nullptr || ws_bytes < need))
    return set_error(WFSP_EWORKSPACE, "conv_wgrad workspace %zu < %zu", ws_bytes, need);
  WgradParams p{};
  p.ca_pad = round_up(c_a, 8);
  p.cb_pad = round_up(c_b, 8);
  __nv_bfloat16* a16 = static_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* b16 = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + align_up(size_t(n_a) * p.ca_pad * 2, 256));
  CastJob ja{a, a16, n_a, c_a, p.ca_pad, n_a_dev}, jb{b, b16, n_b, c_b, p.cb_pad, n_b_dev};
  p.n_a_dev = n_a_dev;
  if (int rc = launch_cast(ja, &jb, st)) return rc;
  p.a = a16; p.n_a = n_a; p.c_a = c_a; p.b = b16; p.n_b = n_b; p.c_b = c_b;
  p.pair_a = pair_a; p.pair_b = pair_b; p.pair_num = pair_num; p.kvol = kvol; p.pitch = pitch; p.dw = d_weight;
  const int n_tiles = (c_b + 255) / 256;
  p.n_tile = round_up((c_b + n_tiles - 1) / n_tiles, 16);
  p.m_tiles = (c_a + kTileM - 1) / kTileM;
  const int tiles = p.m_tiles * n_tiles;
  // pairs per offset that bound the split of the reduction: the caller's hint (graph path, where only
  // capacities are known on the host) or the capacity itself
  int64_t rows = pair_a ? pitch : n_a;
  if (pairs_hint > 0 && pairs_hint < rows) rows = pairs_hint;
  int nsplit = 1;
  if (rows > 0) {
    int64_t want = ceil_div<int64_t>(int64_t(2) * sm_count(), int64_t(tiles) * kvol);
    int64_t maxs = ceil_div<int64_t>(rows, 256);
    nsplit = int(want < 1 ? 1 : (want > maxs ? maxs : want));
    if (int64_t(kvol) * nsplit > 65535) nsplit = 65535 / kvol;
    if (nsplit < 1) nsplit = 1;
  }
  p.nsplit = nsplit;
  p.chunk = round_up(int(ceil_div<int64_t>(rows > 0 ? rows : 1, nsplit)), kSliceK);
  p.use_atomic = (nsplit > 1 || accumulate) ? 1 : 0;
  if (p.use_atomic && !accumulate)
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, size_t(kvol) * c_a * c_b * sizeof(float), st));
  const int b_panels = (p.n_tile + 63) / 64;
  const int stage_bytes = 2 * 8192 + b_panels * 8192;
  p.stages = pick_stages(stage_bytes, 1024);
  const size_t smem = size_t(p.stages) * stage_bytes + 1024;
  dim3 grid(unsigned(tiles), unsigned(kvol * nsplit));
  WFSP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  conv_wgrad_umma_kernel<<<grid, kThreads, smem, st>>>(p);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

}  // namespace wfsp
