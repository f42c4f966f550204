// Shared helpers for libwfsp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "wfsp.h"

namespace wfsp {

// thread-local error text returned by wfsp_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define WFSP_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return wfsp::set_error(WFSP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                               \
  } while (0)

#define WFSP_CHECK_LAUNCH() WFSP_CHECK_CUDA(cudaGetLastError())

#define WFSP_REQUIRE(cond, ...)                                  \
  do {                                                           \
    if (!(cond)) return wfsp::set_error(WFSP_EINVAL, __VA_ARGS__); \
  } while (0)

static inline cudaStream_t as_stream(wfsp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();

// number of kernels libwfsp.so has launched in this process (wfsp_kernel_launches)
void count_launches(int n);

}  // namespace wfsp
