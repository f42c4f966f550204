// Exact-fp32 (WFSP_MATH_FP32) gather-GEMM kernels on CUDA cores.
//
// Forward / dgrad are written output-stationary (SURVEY.md section 7 "hard parts"): a CTA owns a tile of
// destination rows and walks the kernel offsets, gathering the source row nbr[r][k] for each, so
// there is no scatter-add and no atomics; summation order per output element is fixed
// (ascending offset, ascending channel) => bitwise reproducible.  wgrad reduces over the pair
// list of each offset.  These kernels are the tight-tolerance mode and the bring-up reference for
// the tcgen05 path in conv_umma.cu; they replace upstream indiceConv / indiceConvBackward
//  reached from src/models/SPConvBlocks.py:498-502 etc.
#include "common.cuh"

namespace wfsp {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int TM = 4, TN = 4;  // per-thread micro tile; 16x16 threads

// dst[r, n] = bias[n] + sum_k sum_c src[nbr[r,k], c] * W[k][c][n]
// W element (k, c, n) at weight[k*w_ks + c*w_cs + n*w_ns]
__global__ void __launch_bounds__(256) conv_apply_simt_kernel(
    const float* __restrict__ src, int64_t n_src, int c_red, const float* __restrict__ weight, int64_t w_ks,
    int w_cs, int w_ns, const float* __restrict__ bias, const int32_t* __restrict__ nbr, int kvol,
    float* __restrict__ dst, int64_t n_dst, int c_dst, const int32_t* __restrict__ n_src_dev,
    const int32_t* __restrict__ n_dst_dev) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ int s_nbr[BM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = int64_t(blockIdx.x) * BM;
  if (n_src_dev) n_src = *n_src_dev;
  if (n_dst_dev) n_dst = *n_dst_dev;
  if (r0 >= n_dst) return;
  const int n0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < kvol; ++k) {
    int have = 0;
    if (tid < BM) {
      int64_t r = r0 + tid;
      int v = -1;
      if (r < n_dst) v = nbr ? nbr[r * kvol + k] : int(r);
      if (v >= n_src) v = -1;
      s_nbr[tid] = v;
      have = v >= 0;
    }
    if (!__syncthreads_or(have)) continue;  // no row of this tile has a neighbour at offset k
    const float* wk = weight + k * w_ks;
    for (int c0 = 0; c0 < c_red; c0 += BK) {
      // A tile: 64 rows x 16 channels, channel index fastest across threads (coalesced row reads)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int m = (tid >> 4) + 16 * q, kk = tid & 15;
        int row = s_nbr[m];
        int c = c0 + kk;
        As[kk][m] = (row >= 0 && c < c_red) ? src[int64_t(row) * c_red + c] : 0.f;
      }
      // B tile: 16 channels x 64 outputs; thread order follows the unit-stride weight axis
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int kk, n;
        if (w_ns == 1) { n = tid & 63; kk = (tid >> 6) + 4 * q; }
        else           { kk = tid & 15; n = (tid >> 4) + 16 * q; }
        int c = c0 + kk;
        Bs[kk][n] = (c < c_red && n0 + n < c_dst) ? wk[int64_t(c) * w_cs + int64_t(n0 + n) * w_ns] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t r = r0 + ty * TM + i;
    if (r >= n_dst) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < c_dst) dst[r * c_dst + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// d_weight[k][ca][cb] (+)= sum_p a[pa[k,p], ca] * b[pb[k,p], cb]
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(
    const float* __restrict__ a, int64_t n_a, int c_a, const float* __restrict__ b, int64_t n_b, int c_b,
    const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b, const int32_t* __restrict__ pair_num,
    int kvol, int64_t pitch, int nsplit, int64_t chunk, float* __restrict__ d_weight, int use_atomic,
    const int32_t* __restrict__ n_a_dev) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ int s_ia[BK], s_ib[BK];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int a0 = blockIdx.x * BM, b0 = blockIdx.y * BN;
  // NULL lists = identity (1x1 shortcut): one pair per live row
  const int64_t n = pair_num ? int64_t(pair_num[k]) : (n_a_dev ? int64_t(*n_a_dev) : n_a);
  const int64_t p_begin = int64_t(split) * chunk;
  int64_t p_end = p_begin + chunk;
  if (p_end > n) p_end = n;
  if (split > 0 && p_begin >= n) return;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    if (tid < BK) {
      int64_t p = p0 + tid;
      int ia = -1, ib = -1;
      if (p < p_end) {
        ia = pair_a ? pair_a[int64_t(k) * pitch + p] : int(p);
        ib = pair_b ? pair_b[int64_t(k) * pitch + p] : int(p);
      }
      if (ia >= n_a || ib >= n_b) { ia = -1; ib = -1; }
      s_ia[tid] = ia; s_ib[tid] = ib;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int m = tid & 63, kk = (tid >> 6) + 4 * q;
      int ia = s_ia[kk], ib = s_ib[kk];
      As[kk][m] = (ia >= 0 && a0 + m < c_a) ? a[int64_t(ia) * c_a + a0 + m] : 0.f;
      Bs[kk][m] = (ib >= 0 && b0 + m < c_b) ? b[int64_t(ib) * c_b + b0 + m] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dw = d_weight + int64_t(k) * c_a * c_b;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int ca = a0 + ty * TM + i;
    if (ca >= c_a) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int cb = b0 + tx * TN + j;
      if (cb >= c_b) continue;
      if (use_atomic) atomicAdd(&dw[int64_t(ca) * c_b + cb], acc[i][j]);
      else dw[int64_t(ca) * c_b + cb] = acc[i][j];
    }
  }
}

}  // namespace

int conv_apply_simt(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                    const int32_t* n_src_dev, const int32_t* n_dst_dev, cudaStream_t st) {
  if (n_dst == 0) return WFSP_OK;
  dim3 grid(unsigned(ceil_div<int64_t>(n_dst, BM)), unsigned(ceil_div(c_dst, BN)));
  const int64_t w_ks = int64_t(c_red) * c_dst;
  // weight[k] is [c_red, c_dst] (transpose_w == 0) or [c_dst, c_red] (transpose_w == 1)
  const int w_cs = transpose_w ? 1 : c_dst;
  const int w_ns = transpose_w ? c_red : 1;
  conv_apply_simt_kernel<<<grid, 256, 0, st>>>(src, n_src, c_red, weight, w_ks, w_cs, w_ns, bias, nbr, kvol, dst,
                                               n_dst, c_dst, n_src_dev, n_dst_dev);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

int conv_wgrad_simt(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, const int32_t* n_a_dev, cudaStream_t st) {
  const int tiles = ceil_div(c_a, BM) * ceil_div(c_b, BN);
  int nsplit = 1;
  if (pitch > 0) {
    // enough CTAs to fill the machine, at least 256 pairs each
    int64_t want = ceil_div<int64_t>(int64_t(4) * sm_count(), int64_t(tiles) * kvol);
    int64_t maxs = ceil_div<int64_t>(pitch, 256);
    nsplit = int(want < 1 ? 1 : (want > maxs ? maxs : want));
    if (nsplit > 64) nsplit = 64;
    if (int64_t(kvol) * nsplit > 65535) nsplit = 65535 / kvol;
    if (nsplit < 1) nsplit = 1;
  }
  const int64_t chunk = ceil_div<int64_t>(pitch > 0 ? pitch : 1, nsplit);
  const int use_atomic = (nsplit > 1 || accumulate) ? 1 : 0;
  if (use_atomic && !accumulate)
    WFSP_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, size_t(kvol) * c_a * c_b * sizeof(float), st));
  dim3 grid(unsigned(ceil_div(c_a, BM)), unsigned(ceil_div(c_b, BN)), unsigned(kvol * nsplit));
  conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(a, n_a, c_a, b, n_b, c_b, pair_a, pair_b, pair_num, kvol, pitch,
                                               nsplit, chunk, d_weight, use_atomic, n_a_dev);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

}  // namespace wfsp
