// tcgen05 (5th-gen tensor core) gather-GEMM kernels, WFSP_MATH_BF16: bf16 operands, fp32
// accumulation in tensor memory.  They replace upstream indiceConv / indiceConvBackward
// (reference call sites src/models/SPConvBlocks.py:498-502 etc.).
//
// apply kernel (forward, dgrad, inverse forward, inverse dgrad) -- output-stationary implicit GEMM:
//   a CTA owns 128 destination rows.  For every kernel offset k that is active in the tile and
//   every 64-channel slice of the reduction dimension it
//     - gathers the fp32 source rows nbr[r][k] (zeros for -1), converts to bf16 and stores them
//       into shared memory in the 128-byte-swizzled K-major layout UMMA expects (A tile, 16 KB),
//     - copies the matching slice of the pre-transposed bf16 weights (B tile, N x 64),
//     - one thread issues 4 tcgen05.mma (M=128, N=n_tile, K=16) accumulating into TMEM and commits
//       them to the stage's mbarrier, which frees the stage for the gather two/three steps later.
//   The MMAs are asynchronous, so the gather of step i+1 overlaps the tensor work of step i.
//   One epilogue: TMEM -> registers (tcgen05.ld) -> + bias -> coalesced fp32 row stores.
//   No scatter-add, no atomics, fixed summation order.
//
// wgrad kernel: d_weight[k] = A_k^T B_k over the pair list of offset k.  The gathered rows are
//   [pairs][channels] = MN-major operands for UMMA (the reduction index is the pair), so the very
//   same swizzled row layout is used with the MN-major bits set in the instruction descriptor.
//   CTAs split the pair list; partial tiles are reduced with fp32 atomics (red.global.add.f32).
#include "common.cuh"
#include "umma.cuh"

namespace wfsp {
namespace {

using namespace umma;

constexpr int kThreads = 128;
constexpr int kTileM = 128;      // destination rows (apply) / a-channels (wgrad) per CTA
constexpr int kSliceK = 64;      // reduction elements per pipeline stage (one 128 B swizzle row of bf16)
constexpr int kABytes = kTileM * 128;
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 110 * 1024;  // keeps two CTAs resident per SM

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

// load 4 consecutive floats of a row (guarded), VEC = alignment the row pitch guarantees
template <int VEC>
__device__ __forceinline__ void load4(const float* __restrict__ row, int col, int c_max, float (&v)[4]) {
  if (VEC == 4 && col + 3 < c_max) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(row + col));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if (VEC == 2 && col + 3 < c_max) {
    const float2 q0 = __ldg(reinterpret_cast<const float2*>(row + col));
    const float2 q1 = __ldg(reinterpret_cast<const float2*>(row + col + 2));
    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (col + e < c_max) ? __ldg(row + col + e) : 0.f;
  }
}

__device__ __forceinline__ void store_bf16x4(uint8_t* base, uint32_t off, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(base + off) = u;
}

struct ApplyParams {
  const float* src; int64_t n_src; int c_red;
  const __nv_bfloat16* wt; int n_pad, kc_pad;
  const float* bias; const int32_t* nbr; int kvol;
  float* dst; int64_t n_dst; int c_dst;
  int n_tile, stages;
};

template <int VEC>
__global__ void __launch_bounds__(kThreads) conv_apply_umma_kernel(const ApplyParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_free[kMaxStages];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_active[WFSP_MAX_KVOL / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = int64_t(blockIdx.x) * kTileM;
  const int n0 = blockIdx.y * p.n_tile;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kABytes + uint32_t(p.n_tile) * 128u;
  const uint32_t tmem_cols = tmem_cols_pow2(uint32_t(p.n_tile));

  for (int i = tid; i < WFSP_MAX_KVOL / 32; i += kThreads) s_active[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) mbar_init(&bar_free[s], 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // which kernel offsets have at least one neighbour in this tile
  {
    const int64_t r = row0 + tid;
    for (int k = 0; k < p.kvol; ++k) {
      bool v = r < p.n_dst && (p.nbr ? p.nbr[r * p.kvol + k] >= 0 : true);
      unsigned bal = __ballot_sync(0xffffffffu, v);
      if (lane == 0 && bal) atomicOr(&s_active[k >> 5], 1u << (k & 31));
    }
  }
  __syncthreads();

  const int sub = tid & 15;  // 4-float column group inside the 64-wide slice
  const int r8 = tid >> 4;   // this thread covers tile rows r8 + 8*i
  const int num_kb = p.kc_pad / kSliceK;
  const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 0, 0);
  int it = 0;

  for (int kw = 0; kw < (p.kvol + 31) / 32; ++kw) {
    uint32_t mask = s_active[kw];
    while (mask) {
      const int k = kw * 32 + __ffs(mask) - 1;
      mask &= mask - 1;
      int my_rows[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int64_t r = row0 + r8 + 8 * i;
        int v = -1;
        if (r < p.n_dst) v = p.nbr ? __ldg(p.nbr + r * p.kvol + k) : int(r);
        if (v >= p.n_src) v = -1;
        my_rows[i] = v;
      }
      const __nv_bfloat16* wk = p.wt + (int64_t(k) * p.n_pad + n0) * p.kc_pad;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % p.stages;
        if (it >= p.stages) mbar_wait(&bar_free[s], uint32_t((it / p.stages - 1) & 1));
        uint8_t* sa = smem + uint32_t(s) * stage_bytes;
        uint8_t* sb = sa + kABytes;
        const int c0 = kb * kSliceK + sub * 4;
        // A tile: gather + convert
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = my_rows[h * 8 + i];
            if (row >= 0) load4<VEC>(p.src + int64_t(row) * p.c_red, c0, p.c_red, v[i]);
            else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t trow = uint32_t(r8 + 8 * (h * 8 + i));
            store_bf16x4(sa, sw128_offset(trow, uint32_t(sub >> 1)) + uint32_t(sub & 1) * 8u, v[i]);
          }
        }
        // B tile: n_tile rows x 128 B of the prepared weights
        {
          const int c16 = tid & 7;
          for (int n = tid >> 3; n < p.n_tile; n += kThreads / 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(wk + int64_t(n) * p.kc_pad + kb * kSliceK) + c16);
            *reinterpret_cast<uint4*>(sb + sw128_offset(uint32_t(n), uint32_t(c16))) = q;
          }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
          for (int kk = 0; kk < kSliceK / 16; ++kk) {
            const uint64_t adesc = make_desc_sw128(a_addr + kk * 32, 16, 1024);
            const uint64_t bdesc = make_desc_sw128(b_addr + kk * 32, 16, 1024);
            mma_bf16(tmem, adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
          mma_commit(&bar_free[s]);
        }
      }
    }
  }

  if (it > 0) {
    if (tid == 0) mma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    tc_fence_after();
  }
  // epilogue: thread (warp, lane) owns tile row 32*warp + lane
  {
    const int64_t r = row0 + warp * 32 + lane;
    float* out = p.dst + r * p.c_dst;
    const bool vec_ok = (p.c_dst % 4) == 0;
    for (int col = 0; col < p.n_tile; col += 16) {
      uint32_t acc[16];
      if (it > 0) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

struct WgradParams {
  const float* a; int64_t n_a; int c_a;
  const float* b; int64_t n_b; int c_b;
  const int32_t* pair_a; const int32_t* pair_b; const int32_t* pair_num;
  int kvol; int64_t pitch;
  float* dw;
  int n_tile, m_tiles, nsplit, stages, use_atomic;
  int64_t chunk;
};

template <int VEC_A, int VEC_B>
__global__ void __launch_bounds__(kThreads) conv_wgrad_umma_kernel(const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_free[kMaxStages];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.y / p.nsplit, split = blockIdx.y % p.nsplit;
  const int mt = blockIdx.x % p.m_tiles, nt = blockIdx.x / p.m_tiles;
  const int a_c0 = mt * kTileM, b_c0 = nt * p.n_tile;
  const int64_t n_pairs = p.pair_num ? int64_t(p.pair_num[k]) : p.n_a;
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

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) mbar_init(&bar_free[s], 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int sub = tid & 15, r8 = tid >> 4;
  const uint32_t idesc = make_idesc_bf16(kTileM, uint32_t(p.n_tile), 1, 1);
  const int32_t* pa = p.pair_a ? p.pair_a + int64_t(k) * p.pitch : nullptr;
  const int32_t* pb = p.pair_b ? p.pair_b + int64_t(k) * p.pitch : nullptr;

  for (int it = 0; it < iters; ++it) {
    const int s = it % p.stages;
    if (it >= p.stages) mbar_wait(&bar_free[s], uint32_t((it / p.stages - 1) & 1));
    uint8_t* sa = smem + uint32_t(s) * stage_bytes;
    uint8_t* sb = sa + a_bytes;
    int ia[8], ib[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t q = begin + int64_t(it) * kSliceK + r8 + 8 * i;
      int va = -1, vb = -1;
      if (q < end) {
        va = pa ? __ldg(pa + q) : int(q);
        vb = pb ? __ldg(pb + q) : int(q);
      }
      if (va < 0 || vb < 0 || va >= p.n_a || vb >= p.n_b) { va = -1; vb = -1; }
      ia[i] = va; ib[i] = vb;
    }
    // A: [64 pairs][128 channels of a] as two 64-channel swizzled panels
#pragma unroll
    for (int panel = 0; panel < 2; ++panel) {
      float v[8][4];
      const int col = a_c0 + panel * 64 + sub * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (ia[i] >= 0) load4<VEC_A>(p.a + int64_t(ia[i]) * p.c_a, col, p.c_a, v[i]);
        else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        store_bf16x4(sa + panel * 8192, sw128_offset(uint32_t(r8 + 8 * i), uint32_t(sub >> 1)) + uint32_t(sub & 1) * 8u, v[i]);
    }
    // B: [64 pairs][n_tile channels of b]
    for (int panel = 0; panel < b_panels; ++panel) {
      float v[8][4];
      const int col = b_c0 + panel * 64 + sub * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (ib[i] >= 0) load4<VEC_B>(p.b + int64_t(ib[i]) * p.c_b, col, p.c_b, v[i]);
        else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        store_bf16x4(sb + panel * 8192, sw128_offset(uint32_t(r8 + 8 * i), uint32_t(sub >> 1)) + uint32_t(sub & 1) * 8u, v[i]);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
      for (int kk = 0; kk < kSliceK / 16; ++kk) {
        // MN-major: 64-channel panels 8192 B apart (LBO), 8-pair groups 1024 B apart (SBO)
        const uint64_t adesc = make_desc_sw128(a_addr + kk * 2048, 8192, 1024);
        const uint64_t bdesc = make_desc_sw128(b_addr + kk * 2048, 8192, 1024);
        mma_bf16(tmem, adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
      }
      mma_commit(&bar_free[s]);
    }
  }
  if (iters > 0) {
    if (tid == 0) mma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    tc_fence_after();
  }
  {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct ApplyPlan { int n_tiles, n_tile, n_pad, kc_pad; };
ApplyPlan apply_plan(int c_red, int c_dst) {
  ApplyPlan a;
  a.n_tiles = (c_dst + 255) / 256;
  a.n_tile = round_up((c_dst + a.n_tiles - 1) / a.n_tiles, 16);
  a.n_pad = a.n_tiles * a.n_tile;
  a.kc_pad = round_up(c_red, kSliceK);
  return a;
}

inline int vec_of(int c, const void* ptr) {
  if (reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && c % 4 == 0) return 4;
  if (reinterpret_cast<uintptr_t>(ptr) % 8 == 0 && c % 2 == 0) return 2;
  return 1;
}

}  // namespace

size_t conv_apply_umma_workspace(int kvol, int c_red, int c_dst) {
  ApplyPlan a = apply_plan(c_red, c_dst);
  return align_up(size_t(kvol) * a.n_pad * a.kc_pad * sizeof(__nv_bfloat16), 256);
}

int conv_apply_umma(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  if (n_dst == 0) return WFSP_OK;
  ApplyPlan a = apply_plan(c_red, c_dst);
  const size_t need = conv_apply_umma_workspace(kvol, c_red, c_dst);
  if (ws == nullptr || ws_bytes < need) return set_error(WFSP_EWORKSPACE, "conv_apply workspace %zu < %zu", ws_bytes, need);
  __nv_bfloat16* wt = static_cast<__nv_bfloat16*>(ws);
  {
    const int64_t total = int64_t(kvol) * a.n_pad * a.kc_pad;
    int64_t blocks = ceil_div<int64_t>(total, 256);
    if (blocks > int64_t(sm_count()) * 8) blocks = int64_t(sm_count()) * 8;
    prep_weight_kernel<<<unsigned(blocks), 256, 0, st>>>(weight, kvol, c_red, c_dst, transpose_w, wt, a.n_pad, a.kc_pad);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }
  ApplyParams p{src, n_src, c_red, wt, a.n_pad, a.kc_pad, bias, nbr, kvol, dst, n_dst, c_dst, a.n_tile, 2};
  const int stage_bytes = kABytes + a.n_tile * 128;
  int stages = kSmemBudget / stage_bytes;
  p.stages = stages < 2 ? 2 : (stages > kMaxStages ? kMaxStages : stages);
  const size_t smem = size_t(p.stages) * stage_bytes + 1024;
  dim3 grid(unsigned(ceil_div<int64_t>(n_dst, kTileM)), unsigned(a.n_tiles));
  const int vec = vec_of(c_red, src);
  auto launch = [&](auto kern) -> int {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kThreads, smem, st>>>(p);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  };
  if (vec == 4) return launch(conv_apply_umma_kernel<4>);
  if (vec == 2) return launch(conv_apply_umma_kernel<2>);
  return launch(conv_apply_umma_kernel<1>);
}

size_t conv_wgrad_umma_workspace(int, int, int, int64_t) { return 0; }

int conv_wgrad_umma(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, void*, size_t, cudaStream_t st) {
  WgradParams p{};
  p.a = a; p.n_a = n_a; p.c_a = c_a; p.b = b; p.n_b = n_b; p.c_b = c_b;
  p.pair_a = pair_a; p.pair_b = pair_b; p.pair_num = pair_num; p.kvol = kvol; p.pitch = pitch; p.dw = d_weight;
  const int n_tiles = (c_b + 255) / 256;
  p.n_tile = round_up((c_b + n_tiles - 1) / n_tiles, 16);
  p.m_tiles = (c_a + kTileM - 1) / kTileM;
  const int tiles = p.m_tiles * n_tiles;
  const int64_t rows = pair_a ? pitch : n_a;
  int nsplit = 1;
  if (rows > 0) {
    int64_t want = ceil_div<int64_t>(int64_t(2) * sm_count(), int64_t(tiles) * kvol);
    int64_t maxs = ceil_div<int64_t>(rows, 512);
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
  int stages = kSmemBudget / stage_bytes;
  p.stages = stages < 2 ? 2 : (stages > kMaxStages ? kMaxStages : stages);
  const size_t smem = size_t(p.stages) * stage_bytes + 1024;
  dim3 grid(unsigned(tiles), unsigned(kvol * nsplit));
  const int va = vec_of(c_a, a), vb = vec_of(c_b, b);
  auto launch = [&](auto kern) -> int {
    WFSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kThreads, smem, st>>>(p);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  };
  if (va == 4 && vb == 4) return launch(conv_wgrad_umma_kernel<4, 4>);
  if (va >= 2 && vb >= 2) return launch(conv_wgrad_umma_kernel<2, 2>);
  return launch(conv_wgrad_umma_kernel<1, 1>);
}

}  // namespace wfsp
