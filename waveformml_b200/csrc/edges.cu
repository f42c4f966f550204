// Window edges on the GPU: the reference's only native code,
// cffi_window_edges (/root/reference/src/custom_functions/cffi.c:5-37, called from
// src/utils/GraphUtils.py:7-40), which lists for every hit i an optional self loop and, for every later hit j
// of the same contiguous run of equal batch ids with |dx| < n and |dy| < n, the directed edges (i,j), (j,i).
// Output order = the CPU loop's order (bit-exact): element i's edges start at the exclusive prefix sum of the
// per-element edge counts.  It is a rulebook without kernel offsets: count -> scan -> write.
#include "common.cuh"

namespace wfsp {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ bool near_(int64_t a, int64_t b, int64_t n) {
  const int64_t d = a - b;
  return (d < 0 ? -d : d) < n;
}

__global__ void __launch_bounds__(kBlock) edges_count(int64_t n, int64_t num, const int64_t* __restrict__ x,
                                                      const int64_t* __restrict__ y, const int64_t* __restrict__ b,
                                                      int self_loop, int64_t* __restrict__ cnt) {
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= num) return;
  const int64_t bi = b[i], xi = x[i], yi = y[i];
  int64_t c = self_loop ? 1 : 0;
  for (int64_t j = i + 1; j < num && b[j] == bi; ++j)
    if (near_(xi, x[j], n) && near_(yi, y[j], n)) c += 2;
  cnt[i] = c;
}

// exclusive scan of cnt[0..num) by ONE block of 1024 threads walking the array in rounds; total -> *total
__global__ void __launch_bounds__(1024) edges_scan(const int64_t* __restrict__ cnt, int64_t num, int64_t* __restrict__ base,
                                                   int64_t* __restrict__ total) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_run;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int64_t r0 = 0; r0 < num; r0 += 1024) {
    const int64_t i = r0 + tid;
    const int64_t v = i < num ? cnt[i] : 0;
    int64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t off = s_run + incl - v;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (i < num) base[i] = off;
    __syncthreads();
    if (tid == 1023) s_run = off + v;
    __syncthreads();
  }
  if (tid == 0) *total = s_run;
}

__global__ void __launch_bounds__(kBlock) edges_write(int64_t n, int64_t num, const int64_t* __restrict__ x,
                                                      const int64_t* __restrict__ y, const int64_t* __restrict__ b,
                                                      int self_loop, const int64_t* __restrict__ base, int64_t cap,
                                                      int64_t* __restrict__ e1, int64_t* __restrict__ e2) {
  const int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= num) return;
  const int64_t bi = b[i], xi = x[i], yi = y[i];
  int64_t e = base[i];
  if (self_loop) {
    if (e < cap) { e1[e] = i; e2[e] = i; }
    ++e;
  }
  for (int64_t j = i + 1; j < num && b[j] == bi; ++j)
    if (near_(xi, x[j], n) && near_(yi, y[j], n)) {
      if (e + 1 < cap) { e1[e] = i; e2[e] = j; e1[e + 1] = j; e2[e + 1] = i; }
      e += 2;
    }
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" size_t wfsp_window_edges_workspace_bytes(int64_t num_elem) {
  return align_up(size_t(num_elem > 0 ? num_elem : 1) * 2 * sizeof(int64_t), 256);
}

// Two phases so that the caller can size the output exactly: edges1 == NULL -> only *edge_count is produced
// (device scalar); with edges1 / edges2 (capacity edge_cap each) the edges are written as well.
extern "C" int wfsp_window_edges(int64_t n, int64_t num_elem, const int64_t* x, const int64_t* y, const int64_t* b,
                                 int self_loop, int64_t* edges1, int64_t* edges2, int64_t edge_cap, int64_t* edge_count,
                                 void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n >= 0 && num_elem >= 0 && edge_count != nullptr, "bad window-edge arguments");
  cudaStream_t st = as_stream(stream);
  if (num_elem == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(edge_count, 0, sizeof(int64_t), st));
    return WFSP_OK;
  }
  WFSP_REQUIRE(x && y && b, "null coordinates");
  if (workspace == nullptr || workspace_bytes < wfsp_window_edges_workspace_bytes(num_elem))
    return set_error(WFSP_EWORKSPACE, "window-edge workspace too small");
  int64_t* cnt = static_cast<int64_t*>(workspace);
  int64_t* base = cnt + num_elem;
  const unsigned blocks = unsigned(ceil_div<int64_t>(num_elem, kBlock));
  edges_count<<<blocks, kBlock, 0, st>>>(n, num_elem, x, y, b, self_loop, cnt);
  edges_scan<<<1, 1024, 0, st>>>(cnt, num_elem, base, edge_count);
  count_launches(2);
  if (edges1 != nullptr && edges2 != nullptr) {
    edges_write<<<blocks, kBlock, 0, st>>>(n, num_elem, x, y, b, self_loop, base, edge_cap, edges1, edges2);
    count_launches(1);
  }
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
