// C-ABI glue of libwfsp.so: error text, device queries, math-mode dispatch (include/wfsp.h).
#include "common.cuh"

#include <string.h>
#include <atomic>

namespace wfsp {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

void set_force_hash(int v);
void set_small_rows(int v);
void set_force_rblk(int v);
void set_auto_ksplit(int v);
void set_split_stages(int v);
void set_split_wide(int v);
void set_prep_ctas(int v);
void set_bn_stream(int v);
void set_bn_fuse(int v);
void set_trace(unsigned long long* p);

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// conv_simt.cu
int conv_apply_simt(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                    const int32_t* n_src_dev, const int32_t* n_dst_dev, cudaStream_t st);
int conv_wgrad_simt(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, const int32_t* n_a_dev, cudaStream_t st);
// conv_umma.cu
size_t conv_apply_umma_workspace(int kvol, int64_t n_src, int c_red, int c_dst, int split3);
int conv_apply_umma(const float* src, int64_t n_src, int c_red, const float* weight, int transpose_w,
                    const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                    void* ws, size_t ws_bytes, const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                    int split3, cudaStream_t st);
size_t conv_wgrad_umma_workspace(int kvol, int64_t n_a, int c_a, int64_t n_b, int c_b, int64_t pitch, int split3);
int conv_wgrad_umma(const float* a, int64_t n_a, int c_a, const float* b, int64_t n_b, int c_b,
                    const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pitch,
                    float* d_weight, int accumulate, void* ws, size_t ws_bytes, const int32_t* n_a_dev,
                    const int32_t* n_b_dev, int64_t pairs_hint, int split3, cudaStream_t st);

size_t prepared_weight_bytes(int kvol, int c_red, int c_dst);
int prep_weights_batch(const wfsp_prep_job* jobs, int n_jobs, cudaStream_t st);
int cast_rows_bf16(const float* src, int64_t n, const int32_t* n_dev, int c, void* dst16, cudaStream_t st);
int conv_apply_umma_launch(const __nv_bfloat16* act, int64_t n_src, int c_red, const __nv_bfloat16* wt,
                           const float* bias, const int32_t* nbr, int kvol, float* dst, int64_t n_dst, int c_dst,
                           const int32_t* n_src_dev, const int32_t* n_dst_dev, int64_t n_dst_hint,
                           const wfsp_conv_epilogue* ep, cudaStream_t st);
int conv_wgrad_umma_launch(const __nv_bfloat16* a16, int64_t n_a, int c_a, const __nv_bfloat16* b16, int64_t n_b, int c_b,
                           const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num, int kvol,
                           int64_t pitch, float* d_weight, int accumulate, const int32_t* n_a_dev, int64_t pairs_hint,
                           cudaStream_t st);

}  // namespace wfsp

using namespace wfsp;

extern "C" int wfsp_version(void) { return WFSP_VERSION; }

#ifndef WFSP_SOURCE_HASH
#define WFSP_SOURCE_HASH "unknown"
#endif
extern "C" const char* wfsp_source_hash(void) { return WFSP_SOURCE_HASH; }

extern "C" const char* wfsp_last_error(void) { return error_buffer(); }

extern "C" unsigned long long wfsp_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int wfsp_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  WFSP_CHECK_CUDA(cudaGetDevice(&dev));
  WFSP_CHECK_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  WFSP_CHECK_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  WFSP_CHECK_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return WFSP_OK;
}

// test hook: force the open-addressing hash table in the rulebook builder even for small grids
extern "C" int wfsp_set_option(const char* name, int value) {
  if (strcmp(name, "rulebook_force_hash") == 0) { set_force_hash(value); return WFSP_OK; }
  if (strcmp(name, "rulebook_small_rows") == 0) { set_small_rows(value); return WFSP_OK; }
  if (strcmp(name, "apply_row_blocks") == 0) { set_force_rblk(value); return WFSP_OK; }
  if (strcmp(name, "apply_k_split") == 0) { set_auto_ksplit(value); return WFSP_OK; }
  if (strcmp(name, "apply_split_stages") == 0) { set_split_stages(value); return WFSP_OK; }
  if (strcmp(name, "apply_split_wide") == 0) { set_split_wide(value); return WFSP_OK; }
  if (strcmp(name, "prep_ctas") == 0) { set_prep_ctas(value); return WFSP_OK; }
  if (strcmp(name, "bn_stream") == 0) { set_bn_stream(value); return WFSP_OK; }
  if (strcmp(name, "apply_bn_fuse") == 0) { set_bn_fuse(value); return WFSP_OK; }
  return set_error(WFSP_EINVAL, "unknown option %s", name);
}

extern "C" int wfsp_debug_trace(unsigned long long* device_buffer) {
  set_trace(device_buffer);
  return WFSP_OK;
}

extern "C" size_t wfsp_conv_apply_workspace_bytes(int kvol, int64_t n_src, int c_red, int c_dst, int math) {
  if (math == WFSP_MATH_BF16 || math == WFSP_MATH_BF16X3)
    return conv_apply_umma_workspace(kvol, n_src, c_red, c_dst, math == WFSP_MATH_BF16X3);
  return 0;
}

extern "C" int wfsp_conv_apply(const float* src, int64_t n_src, const int32_t* n_src_dev, int c_red,
                               const float* weight, int transpose_w, const float* bias, const int32_t* nbr,
                               int kvol, float* dst, int64_t n_dst, const int32_t* n_dst_dev, int64_t n_dst_hint,
                               int c_dst, int math, void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_src >= 0 && n_dst >= 0 && c_red >= 1 && c_dst >= 1, "bad conv sizes");
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  WFSP_REQUIRE(nbr != nullptr || (kvol == 1 && n_src == n_dst), "identity map needs kvol == 1 and n_src == n_dst");
  if (math == WFSP_MATH_FP32)
    return conv_apply_simt(src, n_src, c_red, weight, transpose_w, bias, nbr, kvol, dst, n_dst, c_dst, n_src_dev,
                           n_dst_dev, as_stream(stream));
  if (math == WFSP_MATH_BF16 || math == WFSP_MATH_BF16X3)
    return conv_apply_umma(src, n_src, c_red, weight, transpose_w, bias, nbr, kvol, dst, n_dst, c_dst, workspace,
                           workspace_bytes, n_src_dev, n_dst_dev, n_dst_hint, math == WFSP_MATH_BF16X3, as_stream(stream));
  return set_error(WFSP_EINVAL, "unknown math mode %d", math);
}

extern "C" size_t wfsp_conv_wgrad_workspace_bytes(int kvol, int64_t n_a, int c_a, int64_t n_b, int c_b,
                                                  int64_t pair_pitch, int math) {
  if (math == WFSP_MATH_BF16 || math == WFSP_MATH_BF16X3)
    return conv_wgrad_umma_workspace(kvol, n_a, c_a, n_b, c_b, pair_pitch, math == WFSP_MATH_BF16X3);
  return 0;
}

extern "C" int wfsp_conv_wgrad(const float* a, int64_t n_a, const int32_t* n_a_dev, int c_a, const float* b,
                               int64_t n_b, const int32_t* n_b_dev, int c_b, const int32_t* pair_a,
                               const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pair_pitch,
                               int64_t pairs_hint, float* d_weight, int accumulate, int math, void* workspace,
                               size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_a >= 0 && n_b >= 0 && c_a >= 1 && c_b >= 1 && pair_pitch >= 0, "bad wgrad sizes");
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  if (math == WFSP_MATH_FP32)
    return conv_wgrad_simt(a, n_a, c_a, b, n_b, c_b, pair_a, pair_b, pair_num, kvol, pair_pitch, d_weight,
                           accumulate, n_a_dev, as_stream(stream));
  if (math == WFSP_MATH_BF16 || math == WFSP_MATH_BF16X3)
    return conv_wgrad_umma(a, n_a, c_a, b, n_b, c_b, pair_a, pair_b, pair_num, kvol, pair_pitch, d_weight,
                           accumulate, workspace, workspace_bytes, n_a_dev, n_b_dev, pairs_hint, math == WFSP_MATH_BF16X3,
                           as_stream(stream));
  return set_error(WFSP_EINVAL, "unknown math mode %d", math);
}

// ---- (6) bf16-resident pipeline -------------------------------------------------------------------
extern "C" size_t wfsp_prepared_weight_bytes(int kvol, int c_red, int c_dst) {
  return prepared_weight_bytes(kvol, c_red, c_dst);
}

extern "C" int wfsp_prep_weights(const wfsp_prep_job* jobs_host, int n_jobs, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_jobs >= 0 && (n_jobs == 0 || jobs_host != nullptr), "bad job list");
  return prep_weights_batch(jobs_host, n_jobs, as_stream(stream));
}

extern "C" int wfsp_cast_rows_bf16(const float* src, int64_t n_rows, const int32_t* n_rows_dev, int c, void* dst_bf16,
                                   wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && c >= 1 && src != nullptr && dst_bf16 != nullptr, "bad cast arguments");
  return cast_rows_bf16(src, n_rows, n_rows_dev, c, dst_bf16, as_stream(stream));
}

extern "C" int wfsp_conv_apply_bf16(const void* src_bf16, int64_t n_src, const int32_t* n_src_dev, int c_red,
                                    const void* weight_prepared, const float* bias, const int32_t* nbr, int kvol,
                                    float* dst, int64_t n_dst, const int32_t* n_dst_dev, int64_t n_dst_hint, int c_dst,
                                    float* bn_partials, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_src >= 0 && n_dst >= 0 && c_red >= 1 && c_dst >= 1, "bad conv sizes");
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  WFSP_REQUIRE(nbr != nullptr || (kvol == 1 && n_src == n_dst), "identity map needs kvol == 1 and n_src == n_dst");
  wfsp_conv_epilogue ep{};
  ep.bn_partials = bn_partials;
  return conv_apply_umma_launch(static_cast<const __nv_bfloat16*>(src_bf16), n_src, c_red,
                                static_cast<const __nv_bfloat16*>(weight_prepared), bias, nbr, kvol, dst, n_dst, c_dst,
                                n_src_dev, n_dst_dev, n_dst_hint, &ep, as_stream(stream));
}

extern "C" int wfsp_conv_apply_bf16_ex(const void* src_bf16, int64_t n_src, const int32_t* n_src_dev, int c_red,
                                       const void* weight_prepared, const float* bias, const int32_t* nbr, int kvol,
                                       float* dst, int64_t n_dst, const int32_t* n_dst_dev, int64_t n_dst_hint, int c_dst,
                                       const wfsp_conv_epilogue* epilogue, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_src >= 0 && n_dst >= 0 && c_red >= 1 && c_dst >= 1, "bad conv sizes");
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  WFSP_REQUIRE(nbr != nullptr || (kvol == 1 && n_src == n_dst), "identity map needs kvol == 1 and n_src == n_dst");
  return conv_apply_umma_launch(static_cast<const __nv_bfloat16*>(src_bf16), n_src, c_red,
                                static_cast<const __nv_bfloat16*>(weight_prepared), bias, nbr, kvol, dst, n_dst, c_dst,
                                n_src_dev, n_dst_dev, n_dst_hint, epilogue, as_stream(stream));
}

extern "C" int wfsp_conv_wgrad_bf16(const void* a_bf16, int64_t n_a, const int32_t* n_a_dev, int c_a, const void* b_bf16,
                                    int64_t n_b, const int32_t* n_b_dev, int c_b, const int32_t* pair_a,
                                    const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pair_pitch,
                                    int64_t pairs_hint, float* d_weight, int accumulate, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_a >= 0 && n_b >= 0 && c_a >= 1 && c_b >= 1 && pair_pitch >= 0, "bad wgrad sizes");
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  (void)n_b_dev;
  return conv_wgrad_umma_launch(static_cast<const __nv_bfloat16*>(a_bf16), n_a, c_a,
                                static_cast<const __nv_bfloat16*>(b_bf16), n_b, c_b, pair_a, pair_b, pair_num, kvol,
                                pair_pitch, d_weight, accumulate, n_a_dev, pairs_hint, as_stream(stream));
}
