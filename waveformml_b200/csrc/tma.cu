// Host side of the TMA helpers (tma.cuh) and a self-test kernel for the gather4 load.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace wfsp {
namespace tma {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
}  // namespace

int make_rows_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int channels, int64_t pitch_elems) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(WFSP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems & 7) != 0)
    return set_error(WFSP_EINVAL, "TMA needs a 16-byte aligned base and row pitch");
  const cuuint64_t dims[2] = {cuuint64_t(channels), cuuint64_t(rows > 0 ? rows : 1)};
  const cuuint64_t strides[1] = {cuuint64_t(pitch_elems) * 2};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WFSP_ECUDA, "cuTensorMapEncodeTiled failed (%d)", int(r));
  return WFSP_OK;
}

namespace {
// loads 128 rows (indices idx[0..127], any value) x 64 channels starting at c0 with 32 gather4 operations and
// writes the tile back row-major, undoing the swizzle with the same formula the UMMA producers assume
__global__ void __launch_bounds__(128) gather4_selftest_kernel(const __grid_constant__ CUtensorMap map, const int32_t* idx,
                                                               int c0, __nv_bfloat16* out /* [128][64] */) {
  __shared__ __align__(1024) uint8_t tile[128 * 128];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&bar)), "r"(128 * 128) : "memory");
  }
  if (tid < 32) {
    gather4(umma::smem_u32(tile) + tid * 512, &map, c0, idx[4 * tid], idx[4 * tid + 1], idx[4 * tid + 2], idx[4 * tid + 3],
            umma::smem_u32(&bar));
  }
  umma::mbar_wait(&bar, 0);
  // thread = row; copy its eight 16-byte chunks from their swizzled positions
  const uint4* src = reinterpret_cast<const uint4*>(tile);
  uint4* dst = reinterpret_cast<uint4*>(out + tid * 64);
  for (int c = 0; c < 8; ++c) dst[c] = src[umma::sw128_offset(uint32_t(tid), uint32_t(c)) >> 4];
}
}  // namespace

}  // namespace tma
}  // namespace wfsp

using namespace wfsp;

// test hook: gathers rows idx[0..127] of the bf16 matrix [rows][pitch] (channels c0 .. c0+63) through TMA
extern "C" int wfsp_selftest_gather4(const void* src_bf16, int64_t rows, int channels, int64_t pitch, const int32_t* idx128,
                                     int c0, void* out_bf16, wfsp_stream_t stream) {
  CUtensorMap map;
  if (int rc = tma::make_rows_map_bf16(&map, src_bf16, rows, channels, pitch)) return rc;
  tma::gather4_selftest_kernel<<<1, 128, 0, as_stream(stream)>>>(map, idx128, c0, static_cast<__nv_bfloat16*>(out_bf16));
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
