// BatchNorm1d(+ReLU) streaming passes for LARGE row counts, fed by bulk copies (cp.async.bulk, the TMA engine without
// a tensor map).  Same arithmetic as the kernels of bn.cu (src/models/SPConvBlocks.py:505-508: nn.BatchNorm1d, nn.ReLU
// between the sparse convolutions); what differs is how the bytes reach the SM.
//
// These passes are pure HBM streams (forward: 4 B read + 2 B written per element; backward: 8 B + 8 B read, 2 B
// written).  With register loads a CTA keeps only (threads x unroll x 8 B) in flight -- 24-64 KB per SM at the
// register budget of the arithmetic -- and measured 0.34-0.58 of the copy bandwidth (profiles/r1_final_timeline_C5).
// Here one elected thread asks the copy engine for whole chunks of rows ([rows, C] is row-major and dense, so a
// chunk of rows is ONE contiguous block): up to three chunks per array are in flight per CTA (~190 KB per SM) at no
// register cost, the 1024 threads only do the arithmetic out of shared memory and the coalesced stores.
//
//   grid   persistent: one CTA per SM (fewer if there are fewer chunks); CTA i takes chunks i, i + grid, ...
//   stage  [x chunk | dy chunk (backward only)]; full[s] mbarrier: one arrival + expect_tx bytes; the stage is handed
//          back to the producer thread by the __syncthreads that ends its consumption
//   tail   a last chunk whose byte count is not a multiple of 16 is copied by the threads themselves
#include "common.cuh"
#include "umma.cuh"

namespace wfsp {
namespace {

using namespace umma;

constexpr int kStThreads = 1024;  // 32 warps: the arithmetic out of shared memory is latency bound per warp, the SM needs all of them
constexpr int kStStages = 3;
constexpr int kStStageBytes = 64 * 1024;  // per stage (all arrays of it)

__device__ __forceinline__ void st_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint32_t st_pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

enum StreamMode { kFwdApply = 0, kBwdPartial = 1, kBwdApply = 2 };

struct StreamParams {
  const float* x; const float* dy;  // dy: backward modes only
  int64_t n_cap; const int32_t* n_dev; int c, chunk_rows;
  const float* gamma; const float* beta; const float* mean; const float* invstd;
  const float* d_gamma; const float* d_beta;  // kBwdApply: the folded sums
  int relu;
  float* out32; __nv_bfloat16* out16;  // kFwdApply: y; kBwdApply: dx
  float* part;                          // kBwdPartial: [gridDim.x][2][c] (sum dy', sum dy' xhat), one entry per CTA
  DropSpec drop;                        // Dropout behind the ReLU (p == 0: none): forward scales y, backward scales dy
};

// the producer's view of chunk q of this CTA: rows [r0, r0 + rows)
__device__ __forceinline__ bool chunk_of(int64_t q, int64_t n, int chunk_rows, int64_t& r0, int& rows) {
  r0 = q * chunk_rows;
  if (r0 >= n) return false;
  rows = int(n - r0 < chunk_rows ? n - r0 : chunk_rows);
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(kStThreads) bn_stream_kernel(const StreamParams p) {
  extern __shared__ uint8_t st_smem_raw[];
  __shared__ uint64_t full[kStStages];
  __shared__ float s_red[2][2][kStThreads];
  constexpr int kArrays = MODE == kFwdApply ? 1 : 2;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 127) & ~uintptr_t(127));
  const int64_t n = p.n_dev ? int64_t(*p.n_dev) : p.n_cap;
  const int c = p.c, tid = threadIdx.x;
  const uint32_t arr_bytes = uint32_t(p.chunk_rows) * c * 4u;           // one array of one stage (full chunk)
  const uint32_t stage_bytes = (kArrays * arr_bytes + 127u) & ~127u;
  const int64_t n_chunks = n > 0 ? (n + p.chunk_rows - 1) / p.chunk_rows : 0;
  // this CTA's chunks: blockIdx.x, + gridDim.x, ...
  const int64_t my_chunks = n_chunks > int64_t(blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int s = 0; s < kStStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // issue the copy of this CTA's i-th chunk into stage i % kStStages (tid 0 only).  A chunk whose byte count is not a
  // multiple of 16 (odd row count x odd c / 2 ...) cannot be a bulk copy: the threads copy it at consumption time.
  auto issue = [&](int64_t i) {
    int64_t r0;
    int rows;
    if (!chunk_of(int64_t(blockIdx.x) + i * gridDim.x, n, p.chunk_rows, r0, rows)) return;
    const uint32_t bytes = uint32_t(rows) * c * 4u;
    const int s = int(i % kStStages);
    const uint32_t bar = smem_u32(&full[s]);
    if ((bytes & 15u) != 0) { st_arrive(bar); return; }  // hand-copied chunk: only the phase has to advance
    const uint32_t dst = smem_u32(smem + size_t(s) * stage_bytes);
    st_expect_tx(bar, kArrays * bytes);
    st_bulk_g2s(dst, p.x + r0 * c, bytes, bar);
    if (kArrays == 2) st_bulk_g2s(dst + arr_bytes, p.dy + r0 * c, bytes, bar);
  };
  if (tid == 0)
    for (int i = 0; i < kStStages && i < my_chunks; ++i) issue(i);

  // thread mapping of the arithmetic: thread x = one pair of adjacent channels, thread y = row lane
  // (c <= 512, so the padded row's pairs fit one pass of the threads)
  const int c_pad = (c + 7) & ~7, ppr = c_pad >> 1;
  const int bx = (ppr + 31) / 32 * 32, by = kStThreads / bx;
  const int tx = tid % bx, ty = tid / bx;
  const bool active = ty < by && tx < ppr;
  const int pc = tx, ch = pc << 1;
  const bool on0 = active && ch < c, on1 = active && ch + 1 < c;
  const float inv_n = n > 0 ? 1.f / float(n) : 0.f;
  uint32_t* out16w = reinterpret_cast<uint32_t*>(p.out16);
  float m0 = 0.f, m1 = 0.f, is0 = 1.f, is1 = 1.f, g0 = 1.f, g1 = 1.f, b0 = 0.f, b1 = 0.f, k0 = 0.f, k1 = 0.f, q0 = 0.f, q1 = 0.f;
  const bool norm = p.mean != nullptr;
  if (norm) {
    if (on0) { m0 = p.mean[ch]; is0 = p.invstd[ch]; if (p.gamma) g0 = p.gamma[ch]; if (p.beta) b0 = p.beta[ch]; }
    if (on1) { m1 = p.mean[ch + 1]; is1 = p.invstd[ch + 1]; if (p.gamma) g1 = p.gamma[ch + 1]; if (p.beta) b1 = p.beta[ch + 1]; }
    if (MODE == kBwdApply) {
      if (on0) { k0 = p.d_beta[ch] * inv_n; q0 = p.d_gamma[ch] * inv_n; }
      if (on1) { k1 = p.d_beta[ch + 1] * inv_n; q1 = p.d_gamma[ch + 1] * inv_n; }
    }
  }
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;  // kBwdPartial: [sum kind][channel of the pair]
  const bool dropping = p.drop.p > 0.f;
  const unsigned long long dkey = dropping ? drop_key(p.drop) : 0ull;
  for (int64_t i = 0; i < my_chunks; ++i) {
    int64_t r0;
    int rows;
    chunk_of(int64_t(blockIdx.x) + i * gridDim.x, n, p.chunk_rows, r0, rows);
    const int s = int(i % kStStages);
    const uint32_t parity = uint32_t((i / kStStages) & 1);
    float* xs = reinterpret_cast<float*>(smem + size_t(s) * stage_bytes);
    float* ds = reinterpret_cast<float*>(smem + size_t(s) * stage_bytes + arr_bytes);
    mbar_wait(&full[s], parity);
    const uint32_t bytes = uint32_t(rows) * c * 4u;
    if ((bytes & 15u) != 0) {  // hand-copied chunk (at most the last one)
      for (int j = tid; j < rows * c; j += kStThreads) {
        xs[j] = p.x[r0 * c + j];
        if (kArrays == 2) ds[j] = p.dy[r0 * c + j];
      }
      __syncthreads();
    }
    if (active) {
#pragma unroll 4
      for (int r = ty; r < rows; r += by) {
        float x0 = 0.f, x1 = 0.f, d0 = 0.f, d1 = 0.f;
        if (on1) {  // c is even: (ch, ch + 1) is one 8-byte aligned pair
          const float2 t = *reinterpret_cast<const float2*>(xs + r * c + ch);
          x0 = t.x; x1 = t.y;
          if (kArrays == 2) { const float2 u = *reinterpret_cast<const float2*>(ds + r * c + ch); d0 = u.x; d1 = u.y; }
        }
        const int64_t gr = r0 + r;
        if (MODE == kFwdApply) {
          float v0 = x0, v1 = x1;
          if (norm) { v0 = (x0 - m0) * is0 * g0 + b0; v1 = (x1 - m1) * is1 * g1 + b1; }
          if (p.relu) { v0 = v0 < 0.f ? 0.f : v0; v1 = v1 < 0.f ? 0.f : v1; }
          if (dropping) {
            v0 *= drop_factor(dkey, p.drop.p, p.drop.scale, (unsigned long long)(gr * c + ch));
            v1 *= drop_factor(dkey, p.drop.p, p.drop.scale, (unsigned long long)(gr * c + ch + 1));
          }
          if (!on1) { v0 = 0.f; v1 = 0.f; }
          if (p.out32 && on1) *reinterpret_cast<float2*>(p.out32 + gr * c + ch) = make_float2(v0, v1);
          if (out16w) out16w[gr * ppr + pc] = st_pack_bf16x2(v0, v1);
        } else {
          const float xh0 = (x0 - m0) * is0, xh1 = (x1 - m1) * is1;
          if (dropping) {
            d0 *= drop_factor(dkey, p.drop.p, p.drop.scale, (unsigned long long)(gr * c + ch));
            d1 *= drop_factor(dkey, p.drop.p, p.drop.scale, (unsigned long long)(gr * c + ch + 1));
          }
          if (norm) {
            if (p.relu && xh0 * g0 + b0 <= 0.f) d0 = 0.f;
            if (p.relu && xh1 * g1 + b1 <= 0.f) d1 = 0.f;
          } else {
            if (p.relu && x0 <= 0.f) d0 = 0.f;
            if (p.relu && x1 <= 0.f) d1 = 0.f;
          }
          if (MODE == kBwdPartial) {
            s00 += d0; s01 += d1; s10 += d0 * xh0; s11 += d1 * xh1;
          } else {
            float v0 = d0, v1 = d1;
            if (norm) { v0 = g0 * is0 * (d0 - k0 - xh0 * q0); v1 = g1 * is1 * (d1 - k1 - xh1 * q1); }
            if (!on1) { v0 = 0.f; v1 = 0.f; }
            if (p.out32 && on1) *reinterpret_cast<float2*>(p.out32 + gr * c + ch) = make_float2(v0, v1);
            if (out16w) out16w[gr * ppr + pc] = st_pack_bf16x2(v0, v1);
          }
        }
      }
    }
    __syncthreads();  // the stage is consumed: the producer refills it with this CTA's chunk i + kStStages
    if (tid == 0 && i + kStStages < my_chunks) issue(i + kStStages);
  }
  if (MODE == kBwdPartial) {
    // combine the row lanes of every channel pair in a fixed order; ONE partial per CTA (zeros if it had no chunk)
    s_red[0][0][tid] = s00; s_red[0][1][tid] = s01; s_red[1][0][tid] = s10; s_red[1][1][tid] = s11;
    __syncthreads();
    if (active && ty == 0) {
      float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
      for (int l = 0; l < by; ++l) {
        t00 += s_red[0][0][l * bx + tx]; t01 += s_red[0][1][l * bx + tx];
        t10 += s_red[1][0][l * bx + tx]; t11 += s_red[1][1][l * bx + tx];
      }
      float* o = p.part + int64_t(blockIdx.x) * 2 * c;
      if (on0) { o[ch] = t00; o[c + ch] = t10; }
      if (on1) { o[ch + 1] = t01; o[c + ch + 1] = t11; }
    }
  }
}

}  // namespace

// rows per chunk: as many as fit kStStageBytes for `arrays` arrays, even (so that every full chunk of an even-c
// tensor is a whole number of 16-byte units), at least 2
int bn_stream_chunk_rows(int c, int arrays) {
  int rows = kStStageBytes / (arrays * c * 4);
  rows &= ~1;
  if (rows > 256) rows = 256;
  return rows < 2 ? 0 : rows;
}

int bn_stream_grid(int64_t n_rows, int chunk_rows) {
  int64_t chunks = ceil_div<int64_t>(n_rows > 0 ? n_rows : 1, chunk_rows);
  const int64_t cap = sm_count();
  return int(chunks < cap ? chunks : cap);
}

// usable when full chunks are 16-byte multiples at 16-byte aligned addresses
bool bn_stream_ok(int c, int arrays, const void* x, const void* dy) {
  if ((c & 1) != 0 || c < 2 || c > 512) return false;
  if (bn_stream_chunk_rows(c, arrays) == 0) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (dy != nullptr && (reinterpret_cast<uintptr_t>(dy) & 15) != 0)) return false;
  return true;
}

template <int MODE>
static int launch_stream(StreamParams p, int arrays, int grid, cudaStream_t st) {
  const size_t arr_bytes = size_t(p.chunk_rows) * p.c * 4;
  const size_t stage_bytes = (arrays * arr_bytes + 127) & ~size_t(127);
  const size_t smem = kStStages * stage_bytes + 256;
  WFSP_CHECK_CUDA(cudaFuncSetAttribute(bn_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  bn_stream_kernel<MODE><<<unsigned(grid), kStThreads, smem, st>>>(p);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

int bn_stream_fwd_apply(const float* x, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma, const float* beta,
                        const float* mean, const float* invstd, int relu, float* y, void* y16, const DropSpec& drop,
                        cudaStream_t st) {
  StreamParams p{};
  p.drop = drop;
  p.x = x; p.n_cap = n_rows; p.n_dev = n_dev; p.c = c; p.chunk_rows = bn_stream_chunk_rows(c, 1);
  p.gamma = gamma; p.beta = beta; p.mean = mean; p.invstd = invstd; p.relu = relu;
  p.out32 = y; p.out16 = static_cast<__nv_bfloat16*>(y16);
  return launch_stream<kFwdApply>(p, 1, bn_stream_grid(n_rows, p.chunk_rows), st);
}

// part: [grid][2][c] floats, grid = bn_stream_grid(n_rows, bn_stream_chunk_rows(c, 2)); returns the grid through *n_part
int bn_stream_bwd_partial(const float* x, const float* dy, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma,
                          const float* beta, const float* mean, const float* invstd, int relu, float* part, int* n_part,
                          const DropSpec& drop, cudaStream_t st) {
  StreamParams p{};
  p.drop = drop;
  p.x = x; p.dy = dy; p.n_cap = n_rows; p.n_dev = n_dev; p.c = c; p.chunk_rows = bn_stream_chunk_rows(c, 2);
  p.gamma = gamma; p.beta = beta; p.mean = mean; p.invstd = invstd; p.relu = relu; p.part = part;
  const int grid = bn_stream_grid(n_rows, p.chunk_rows);
  *n_part = grid;
  return launch_stream<kBwdPartial>(p, 2, grid, st);
}

int bn_stream_bwd_apply(const float* x, const float* dy, int64_t n_rows, const int32_t* n_dev, int c, const float* gamma,
                        const float* beta, const float* mean, const float* invstd, const float* d_gamma, const float* d_beta,
                        int relu, float* dx, void* dx16, const DropSpec& drop, cudaStream_t st) {
  StreamParams p{};
  p.drop = drop;
  p.x = x; p.dy = dy; p.n_cap = n_rows; p.n_dev = n_dev; p.c = c; p.chunk_rows = bn_stream_chunk_rows(c, 2);
  p.gamma = gamma; p.beta = beta; p.mean = mean; p.invstd = invstd; p.d_gamma = d_gamma; p.d_beta = d_beta; p.relu = relu;
  p.out32 = dx; p.out16 = static_cast<__nv_bfloat16*>(dx16);
  return launch_stream<kBwdApply>(p, 2, bn_stream_grid(n_rows, p.chunk_rows), st);
}

}  // namespace wfsp
