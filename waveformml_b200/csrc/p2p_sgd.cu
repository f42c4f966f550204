// Data-parallel gradient exchange fused with the optimiser over NVLink peer memory (one process per GPU, events sharded
// by rank).  The reference does this with Lightning's DDP: NCCL all-reduce of gradient buckets, then
// torch.optim.SGD on every rank (src/utils/util.py:233-236, config/examples/GEP.json:56-68).  Here the flat gradient
// and parameter buffers of all ranks live in peer-mapped (symmetric) memory and ONE kernel does
//
//     reduce-scatter   rank r sums shard r of every rank's gradient, in rank order, reading its peers over NVLink
//     SGD              momentum / Nesterov update of shard r (the 1 / world mean folded in), momentum kept by rank r
//     all-gather       the updated parameter shard is stored straight into every rank's parameter buffer
//
// so a step costs one pass over 1/world of the buffers per GPU plus two cross-GPU barriers, instead of an all-reduce
// of the whole 4.2 MB buffer followed by the optimiser on all of it.  Every parameter is computed by exactly one rank
// and copied to the others: all ranks end bit-identical, whatever the timing.
//
// Cross-GPU barriers are flags in peer memory: rank r stores the step's epoch into slot r of every peer's flag array
// (release, system scope); a waiter polls its OWN array (acquire, system scope).  Barrier 1 (flags[0..world)): every
// rank's backward pass is complete -> gradients may be read.  Barrier 2 (flags[world..2 world)): every rank has
// finished reading gradients and writing parameters -> the next step may overwrite / read them.  Polls are bounded
// (trap instead of hanging the box if a rank died).
#include "common.cuh"

namespace wfsp {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread t < world waits until flags[t] has reached `epoch` (flags only ever grow; wrap-around after 4e9 steps)
__device__ __forceinline__ void wait_flags(const unsigned* flags, int world, unsigned epoch) {
  if (int(threadIdx.x) < world) {
    const long long t0 = clock64();
    while (int(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) __trap();
    }
  }
}

struct P2PParams {
  float* params; const float* grads; float* momentum_buf;
  const float* const* peer_grads; float* const* peer_params; unsigned* const* peer_flags;
  unsigned* flags;     // this rank's flag array [2 * world]
  unsigned* epoch;     // device step counter (local)
  unsigned* ticket;    // zero-initialised CTA counter (left zero)
  unsigned* pending;   // set once this rank has arrived at barrier 2; cleared by the wait kernel
  int64_t n;
  int rank, world;
  float lr, momentum, weight_decay, grad_scale;
  int nesterov;
};

constexpr int kP2PThreads = 512;

__global__ void __launch_bounds__(kP2PThreads) p2p_sgd_kernel(const P2PParams p) {
  __shared__ unsigned s_ticket;
  const unsigned epoch = *p.epoch + 1u;
  const int tid = threadIdx.x;
  // barrier 1: announce "my gradients are complete" (stream order guarantees it) to every rank, wait for all of them
  if (blockIdx.x == 0 && tid < p.world) st_release_sys(p.peer_flags[tid] + p.rank, epoch);
  wait_flags(p.flags, p.world, epoch);
  __syncthreads();
  // shard of this rank, in float4 units
  const int64_t n4 = (p.n + 3) / 4;
  const int64_t per = (n4 + p.world - 1) / p.world;
  const int64_t lo = int64_t(p.rank) * per, hi = lo + per < n4 ? lo + per : n4;
  for (int64_t i = lo + int64_t(blockIdx.x) * kP2PThreads + tid; i < hi; i += int64_t(gridDim.x) * kP2PThreads) {
    const int64_t e0 = i * 4;
    const bool full = e0 + 4 <= p.n;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (full) {
      float4 v[8];
      for (int q0 = 0; q0 < p.world; q0 += 8) {  // eight peers' loads in flight, added in rank order
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (q0 + u < p.world) v[u] = __ldcv(reinterpret_cast<const float4*>(p.peer_grads[q0 + u] + e0));
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (q0 + u < p.world) { g[0] += v[u].x; g[1] += v[u].y; g[2] += v[u].z; g[3] += v[u].w; }
      }
    } else {
      for (int q = 0; q < p.world; ++q)
        for (int e = 0; e < 4; ++e)
          if (e0 + e < p.n) g[e] += __ldcv(p.peer_grads[q] + e0 + e);
    }
    float w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e0 + e >= p.n) { w[e] = 0.f; continue; }
      const float pv = p.params[e0 + e];
      float gv = g[e] * p.grad_scale + p.weight_decay * pv;
      float step = gv;
      if (p.momentum != 0.f) {
        const float b = p.momentum * p.momentum_buf[e0 + e] + gv;
        p.momentum_buf[e0 + e] = b;
        step = p.nesterov ? gv + p.momentum * b : b;
      }
      w[e] = pv - p.lr * step;
    }
    for (int q = 0; q < p.world; ++q) {
      if (full) *reinterpret_cast<float4*>(p.peer_params[q] + e0) = make_float4(w[0], w[1], w[2], w[3]);
      else
        for (int e = 0; e < 4; ++e)
          if (e0 + e < p.n) p.peer_params[q][e0 + e] = w[e];
    }
  }
  // barrier 2, arrival: once EVERY CTA of this rank has issued its peer stores (and finished reading peer gradients).
  // One thread fences for its CTA: the block barrier orders the other threads' stores before that fence (fence
  // cumulativity), so 75 000 system-scope fences are not needed.
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    s_ticket = atomicAdd(p.ticket, 1u);
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (tid == 0) {
    __threadfence_system();
    *p.ticket = 0u;
    *p.pending = 1u;
  }
  __syncthreads();
  if (tid < p.world) st_release_sys(p.peer_flags[tid] + p.world + p.rank, epoch);
}

// barrier 2, wait: every rank's parameter stores have landed here and nobody reads this rank's gradients any more
// (idempotent: without an exchange since the last wait it returns at once, so a captured step may always begin with it)
__global__ void __launch_bounds__(32) p2p_wait_kernel(const unsigned* flags, int world, unsigned* state) {
  if (state[2] == 0u) return;
  const unsigned e = state[0] + 1u;
  wait_flags(flags + world, world, e);
  __syncwarp();
  if (threadIdx.x == 0) { state[0] = e; state[2] = 0u; }
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" int wfsp_sgd_step_p2p(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                                 int nesterov, float weight_decay, float grad_scale, const void* peer_grads_dev,
                                 const void* peer_params_dev, const void* peer_flags_dev, void* flags, void* state, int rank,
                                 int world, int wait_now, wfsp_stream_t stream) {
  WFSP_REQUIRE(n >= 0 && params && grads && peer_grads_dev && peer_params_dev && peer_flags_dev && flags && state,
               "bad peer-memory optimiser arguments");
  WFSP_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "bad rank / world size");
  WFSP_REQUIRE(momentum == 0.f || momentum_buf != nullptr, "momentum needs a buffer");
  WFSP_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0,
               "flat buffers must be 16-byte aligned");
  P2PParams p{};
  p.params = params; p.grads = grads; p.momentum_buf = momentum_buf;
  p.peer_grads = static_cast<const float* const*>(peer_grads_dev);
  p.peer_params = static_cast<float* const*>(peer_params_dev);
  p.peer_flags = static_cast<unsigned* const*>(peer_flags_dev);
  p.flags = static_cast<unsigned*>(flags);
  p.epoch = static_cast<unsigned*>(state);
  p.ticket = static_cast<unsigned*>(state) + 1;
  p.pending = static_cast<unsigned*>(state) + 2;
  p.n = n; p.rank = rank; p.world = world;
  p.lr = lr; p.momentum = momentum; p.weight_decay = weight_decay; p.grad_scale = grad_scale; p.nesterov = nesterov;
  const int64_t n4 = (n + 3) / 4, per = (n4 + world - 1) / world;
  // two 16-byte elements per thread: few CTAs (each polls the flags and fences once), enough loads in flight
  int64_t blocks = ceil_div<int64_t>(per > 0 ? per : 1, kP2PThreads * 2);
  const int64_t cap = sm_count();
  if (blocks > cap) blocks = cap;
  cudaStream_t st = as_stream(stream);
  p2p_sgd_kernel<<<unsigned(blocks), kP2PThreads, 0, st>>>(p);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  if (wait_now) return wfsp_sgd_p2p_wait(flags, state, world, stream);
  return WFSP_OK;
}

extern "C" int wfsp_sgd_p2p_wait(void* flags, void* state, int world, wfsp_stream_t stream) {
  WFSP_REQUIRE(flags && state && world >= 1 && world <= 32, "bad peer-memory barrier arguments");
  p2p_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<unsigned*>(flags), world, static_cast<unsigned*>(state));
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}
