// TMA helpers: tensor-map construction on the host (driver entry point fetched through the runtime, no
// libcuda link) and the gather4 bulk-tensor load that fetches four arbitrary rows of a [rows][channels]
// matrix into four consecutive 128-byte-swizzled shared-memory rows with ONE instruction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wfsp {
namespace tma {

// tensor map of a bf16 row-major matrix [rows][pitch_elems] (pitch a multiple of 8 elements = 16 bytes),
// box = 64 channels x 1 row (gather4 fetches four such rows), 128-byte swizzle, out-of-bounds reads as zero
int make_rows_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int channels, int64_t pitch_elems);

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// rows r0..r3 (any order, any value: rows outside the tensor read as zero), channels [c0, c0 + 64) ->
// 4 x 128 B at dst (shared, 128-byte swizzle: dst must be a multiple of 512 B inside a 1024-B-aligned tile);
// completes 512 bytes on `bar`
__device__ __forceinline__ void gather4(uint32_t dst_smem, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                        uint32_t bar_smem) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_smem)
      : "memory");
}

}  // namespace tma
}  // namespace wfsp
