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

// Training-mode Dropout of the blocks conv . BatchNorm . ReLU . Dropout (src/models/SPConvBlocks.py:375-376, 509-510),
// fused into the kernels that write / read the block's output: the keep / drop decision of element (row, channel) is
// a counter-based hash of (seed, device step counter, layer salt, row * c + channel), so the backward pass REGENERATES
// the forward mask instead of storing it.  p == 0 switches it off.
struct DropSpec {
  float p, scale;
  unsigned long long seed;
  const long long* step;  // device counter (graph path: advances every replay), may be NULL
  unsigned salt;
};
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long drop_key(const DropSpec& d) {
  unsigned long long k = d.seed ^ ((unsigned long long)d.salt << 32);
  if (d.step) k ^= (unsigned long long)(*d.step) * 0x9E3779B97F4A7C15ull;
  return k;
}
// 1 / (1 - p) if element idx is kept, else 0 (splitmix64 finaliser of key + idx * golden ratio; 24 random bits)
__device__ __forceinline__ float drop_factor(unsigned long long key, float p, float scale, unsigned long long idx) {
  unsigned long long z = key + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return float(unsigned(z >> 40)) * (1.f / 16777216.f) < p ? 0.f : scale;
}
#endif
inline DropSpec make_drop(const wfsp_dropout* d) {
  DropSpec s{};
  if (d != nullptr && d->p > 0.f) {
    s.p = d->p;
    s.scale = 1.f / (1.f - d->p);
    s.seed = d->seed;
    s.step = reinterpret_cast<const long long*>(d->step_dev);
    s.salt = d->salt;
  }
  return s;
}

}  // namespace wfsp
