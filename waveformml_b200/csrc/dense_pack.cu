// Batcher (pack) and dense scatter / gather kernels: pure copy / convert work, HBM-bound.
//
//  * wfsp_batch_pack   : collate_fn (src/engineering/PSDDataModule.py:10-20) + int16 -> float
//                        normalisation (src/datasets/HDF5Dataset.py:345-346) + batch-first permute
//                        (src/models/SPConvNet.py:63-64) in one pass: 2 B read + 4 B written / sample.
//  * wfsp_to_dense     : SparseConvTensor.dense() / spconv.ToDense (src/models/SPConvBlocks.py:81):
//                        [N,C] rows -> [B,C,H,W]; a 32x32 shared-memory transpose so that both the
//                        row reads (channel-contiguous) and the NCHW writes (cell-contiguous) are
//                        coalesced; every dense element is written exactly once (no memset pass).
//  * wfsp_to_dense_bwd : the gather back, same transpose in the other direction.
#include "common.cuh"

namespace wfsp {
namespace {

__global__ void __launch_bounds__(256) pack_indices_kernel(const int32_t* __restrict__ coords, int64_t n,
                                                           const int32_t* __restrict__ n_dev,
                                                           const int64_t* __restrict__ item_rows,
                                                           const int64_t* __restrict__ item_offset, int64_t n_items,
                                                           int32_t* __restrict__ indices) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (j >= n) return;
  // item of row j: last it with item_rows[it] <= j
  int64_t lo = 0, hi = n_items - 1;
  while (lo < hi) {
    int64_t mid = (lo + hi + 1) >> 1;
    if (item_rows[mid] <= j) lo = mid; else hi = mid - 1;
  }
  int32_t off = n_items > 0 ? int32_t(item_offset[lo]) : 0;
  indices[3 * j + 0] = coords[3 * j + 2] + off;
  indices[3 * j + 1] = coords[3 * j + 0];
  indices[3 * j + 2] = coords[3 * j + 1];
}

template <typename In>
__device__ __forceinline__ float to_f32(In v) { return float(v); }

template <typename In, typename Out>
__global__ void __launch_bounds__(256) pack_feats_kernel(const In* __restrict__ wave, int64_t n,
                                                         const int32_t* __restrict__ n_dev, int c, float scale,
                                                         Out* __restrict__ feats, int64_t pitch) {
  // one thread per 4 consecutive channels (c % 4 == 0 fast path), grid-stride
  if (n_dev) n = *n_dev;
  const int c4 = c >> 2;
  const int64_t total = n * c4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t row = i / c4;
    int col = int(i - row * c4) << 2;
    const In* s = wave + row * c + col;
    float v0, v1, v2, v3;
    if (sizeof(In) == 2) {
      const short4 q = *reinterpret_cast<const short4*>(s);
      v0 = float(q.x); v1 = float(q.y); v2 = float(q.z); v3 = float(q.w);
    } else {
      const float4 q = *reinterpret_cast<const float4*>(s);
      v0 = q.x; v1 = q.y; v2 = q.z; v3 = q.w;
    }
    v0 *= scale; v1 *= scale; v2 *= scale; v3 *= scale;
    Out* d = feats + row * pitch + col;
    if (sizeof(Out) == 4) {
      *reinterpret_cast<float4*>(d) = make_float4(v0, v1, v2, v3);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(d) = u;
      if (col + 4 == c)  // the row's last chunk also zeroes the pitch padding (bf16 operand format: pitch = c rounded up to 8)
        for (int64_t pc = c; pc < pitch; ++pc) feats[row * pitch + pc] = Out(0.f);
    }
  }
}

template <typename In, typename Out>
__global__ void __launch_bounds__(256) pack_feats_scalar_kernel(const In* __restrict__ wave, int64_t n,
                                                                const int32_t* __restrict__ n_dev, int c,
                                                                float scale, Out* __restrict__ feats, int64_t pitch) {
  if (n_dev) n = *n_dev;
  const int64_t total = n * c;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t row = i / c;
    int col = int(i - row * c);
    float v = float(wave[i]) * scale;
    if (sizeof(Out) == 4) reinterpret_cast<float*>(feats)[row * pitch + col] = v;
    else reinterpret_cast<__nv_bfloat16*>(feats)[row * pitch + col] = __float2bfloat16_rn(v);
  }
}

template <typename In, typename Out>
int launch_pack_feats(const void* wave, int64_t n, const int32_t* n_dev, int c, float scale, void* feats,
                      int64_t pitch, cudaStream_t st) {
  if (n == 0 || c == 0) return WFSP_OK;
  const bool vec = (c % 4 == 0) && (pitch % 4 == 0) && (reinterpret_cast<uintptr_t>(wave) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(feats) % 16 == 0);
  const int64_t work = vec ? n * (c / 4) : n * int64_t(c);
  int64_t blocks = ceil_div<int64_t>(work, 256);
  const int64_t cap = int64_t(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    pack_feats_kernel<In, Out><<<unsigned(blocks), 256, 0, st>>>(static_cast<const In*>(wave), n, n_dev, c, scale,
                                                                 static_cast<Out*>(feats), pitch);
  else
    pack_feats_scalar_kernel<In, Out><<<unsigned(blocks), 256, 0, st>>>(static_cast<const In*>(wave), n, n_dev, c, scale,
                                                                        static_cast<Out*>(feats), pitch);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

// ---- dense ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_mark_kernel(const int32_t* __restrict__ indices, int64_t n,
                                                         const int32_t* __restrict__ n_dev, int batch, int h, int w,
                                                         int32_t* __restrict__ cell_table) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (j >= n) return;
  int b = indices[3 * j], x = indices[3 * j + 1], y = indices[3 * j + 2];
  if (b < 0 || b >= batch || x < 0 || x >= h || y < 0 || y >= w) return;
  atomicMax(&cell_table[(int64_t(b) * h + x) * w + y], int(j));  // last row wins, as index assignment does
}

// grid: (ceil(cells/32), ceil(C/32)), block (32, 8)
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ feats, int c,
                                                        const int32_t* __restrict__ cell_table, int64_t cells,
                                                        int hw, float* __restrict__ dense) {
  __shared__ float tile[32][33];
  __shared__ int s_row[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t cell0 = int64_t(blockIdx.x) * 32;
  const int ch0 = blockIdx.y * 32;
  if (ty == 0) s_row[tx] = (cell0 + tx < cells) ? cell_table[cell0 + tx] : -1;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int i = ty + 8 * q;  // cell within tile
    int row = s_row[i];
    tile[i][tx] = (row >= 0 && ch0 + tx < c) ? feats[int64_t(row) * c + ch0 + tx] : 0.f;
  }
  __syncthreads();
  const int64_t cell = cell0 + tx;
  if (cell >= cells) return;
  const int64_t b = cell / hw;
  const int xy = int(cell - b * hw);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int ch = ch0 + ty + 8 * q;
    if (ch < c) dense[(b * c + ch) * hw + xy] = tile[tx][ty + 8 * q];
  }
}

// grid: (ceil(N/32), ceil(C/32)), block (32, 8)
__global__ void __launch_bounds__(256) dense_bwd_kernel(const float* __restrict__ d_dense,
                                                        const int32_t* __restrict__ indices, int64_t n,
                                                        const int32_t* __restrict__ n_dev, int c, int batch, int h,
                                                        int w, float* __restrict__ d_feats) {
  __shared__ float tile[32][33];
  __shared__ int64_t s_base[32];
  __shared__ int s_xy[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t row0 = int64_t(blockIdx.x) * 32;
  const int ch0 = blockIdx.y * 32;
  const int hw = h * w;
  if (n_dev) n = *n_dev;
  if (row0 >= n) return;
  if (ty == 0) {
    int64_t j = row0 + tx;
    int64_t base = -1;
    int xy = 0;
    if (j < n) {
      int b = indices[3 * j], x = indices[3 * j + 1], y = indices[3 * j + 2];
      if (b >= 0 && b < batch && x >= 0 && x < h && y >= 0 && y < w) { base = int64_t(b) * c * hw; xy = x * w + y; }
    }
    s_base[tx] = base;
    s_xy[tx] = xy;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int ch = ch0 + ty + 8 * q;  // tx indexes the row: neighbouring rows are neighbouring cells
    int64_t base = s_base[tx];
    tile[ty + 8 * q][tx] = (base >= 0 && ch < c) ? d_dense[base + int64_t(ch) * hw + s_xy[tx]] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int i = ty + 8 * q;  // row within tile
    int64_t j = row0 + i;
    if (j < n && ch0 + tx < c) d_feats[j * c + ch0 + tx] = tile[tx][i];
  }
}

// ---- masked L1 segment loss of the z / energy regression models (src/engineering/LitBase.py:124-174,
// _calc_segment_loss: ones-mask and target densified through SparseConvTensor.dense(), l1_loss(mask * prediction,
// target, "sum") / N).  Inactive cells contribute |0 - 0|, so the loss is the row-wise L1 between the dense prediction
// at every hit's cell and that hit's target: one gather pass instead of three ToDense + ~15 element-wise kernels.
// pred [B, C, H, W]; target [n, Ct] with Ct = C or 1 (broadcast over channels, as the dense subtraction would).
constexpr int kSegRows = 256;

__global__ void __launch_bounds__(256) seg_l1_partial(const float* __restrict__ pred, const int32_t* __restrict__ indices,
                                                      const float* __restrict__ target, int64_t n_cap,
                                                      const int32_t* __restrict__ n_dev, int C, int Ct, int batch, int h, int w,
                                                      float* __restrict__ part) {
  __shared__ float s_red[8];
  const int64_t n = n_dev ? int64_t(*n_dev) : n_cap;
  const int64_t j = int64_t(blockIdx.x) * kSegRows + threadIdx.x;
  float s = 0.f;
  if (j < n) {
    const int b = indices[3 * j], x = indices[3 * j + 1], y = indices[3 * j + 2];
    if (b >= 0 && b < batch && x >= 0 && x < h && y >= 0 && y < w)
      for (int c = 0; c < C; ++c)
        s += fabsf(pred[((int64_t(b) * C + c) * h + x) * w + y] - target[j * Ct + (Ct == 1 ? 0 : c)]);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_red[k];
    part[blockIdx.x] = t;
  }
}

// one CTA: chunk partials added in chunk order (strided over 256 lanes, lanes combined in lane order), / N
__global__ void __launch_bounds__(256) seg_l1_finish(const float* __restrict__ part, int64_t n_cap, const int32_t* __restrict__ n_dev,
                                                     float* __restrict__ loss) {
  __shared__ double s_red[256];
  const int64_t n = n_dev ? int64_t(*n_dev) : n_cap;
  const int64_t chunks = (n + kSegRows - 1) / kSegRows;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < chunks; i += 256) s += part[i];
  s_red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 256; ++k) t += s_red[k];
    *loss = n > 0 ? float(t / double(n)) : 0.f;
  }
}

// d_pred[b, c, x, y] = sign(pred - target) * grad_out / N at every hit's cell (d_pred is zeroed by the caller)
__global__ void __launch_bounds__(256) seg_l1_bwd(const float* __restrict__ pred, const int32_t* __restrict__ indices,
                                                  const float* __restrict__ target, int64_t n_cap,
                                                  const int32_t* __restrict__ n_dev, int C, int Ct, int batch, int h, int w,
                                                  const float* __restrict__ grad_out, float* __restrict__ d_pred) {
  const int64_t n = n_dev ? int64_t(*n_dev) : n_cap;
  const int64_t j = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (j >= n) return;
  const float g = (grad_out ? *grad_out : 1.f) / float(n);
  const int b = indices[3 * j], x = indices[3 * j + 1], y = indices[3 * j + 2];
  if (b < 0 || b >= batch || x < 0 || x >= h || y < 0 || y >= w) return;
  for (int c = 0; c < C; ++c) {
    const int64_t cell = ((int64_t(b) * C + c) * h + x) * w + y;
    const float d = pred[cell] - target[j * Ct + (Ct == 1 ? 0 : c)];
    d_pred[cell] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
  }
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" size_t wfsp_segment_l1_workspace_bytes(int64_t n_rows) {
  return align_up(size_t(ceil_div<int64_t>(n_rows > 0 ? n_rows : 1, kSegRows)) * sizeof(float), 256);
}

extern "C" int wfsp_segment_l1_fwd(const float* pred, const int32_t* indices, const float* target, int64_t n_rows,
                                   const int32_t* n_rows_dev, int n_chan, int target_chan, int batch, int h, int w, float* loss,
                                   void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && n_chan >= 1 && (target_chan == n_chan || target_chan == 1) && batch >= 0 && h > 0 && w > 0,
               "bad segment-loss sizes");
  WFSP_REQUIRE(pred && loss && (n_rows == 0 || (indices && target)), "null argument");
  if (workspace == nullptr || workspace_bytes < wfsp_segment_l1_workspace_bytes(n_rows))
    return set_error(WFSP_EWORKSPACE, "segment-loss workspace too small");
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(workspace);
  if (n_rows > 0)
    seg_l1_partial<<<unsigned(ceil_div<int64_t>(n_rows, kSegRows)), 256, 0, st>>>(pred, indices, target, n_rows, n_rows_dev, n_chan,
                                                                                target_chan, batch, h, w, part);
  seg_l1_finish<<<1, 256, 0, st>>>(part, n_rows, n_rows_dev, loss);
  count_launches(2);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_segment_l1_bwd(const float* pred, const int32_t* indices, const float* target, int64_t n_rows,
                                   const int32_t* n_rows_dev, int n_chan, int target_chan, int batch, int h, int w,
                                   const float* grad_out, float* d_pred, wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && n_chan >= 1 && (target_chan == n_chan || target_chan == 1), "bad segment-loss sizes");
  if (n_rows == 0) return WFSP_OK;
  seg_l1_bwd<<<unsigned(ceil_div<int64_t>(n_rows, 256)), 256, 0, as_stream(stream)>>>(pred, indices, target, n_rows, n_rows_dev,
                                                                                    n_chan, target_chan, batch, h, w, grad_out, d_pred);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_batch_pack(const int32_t* coords_xye, const void* wave, int wave_dtype, int64_t n_rows,
                               const int32_t* n_rows_dev, int n_chan, const int64_t* item_rows, const int64_t* item_offset, int64_t n_items,
                               float scale, int32_t* indices_bxy, void* feats, int feats_dtype, int64_t feats_pitch,
                               wfsp_stream_t stream) {
  WFSP_REQUIRE(n_rows >= 0 && n_chan >= 0 && feats_pitch >= n_chan, "bad pack sizes");
  WFSP_REQUIRE(wave_dtype == WFSP_I16 || wave_dtype == WFSP_F32, "wave dtype must be int16 or f32");
  WFSP_REQUIRE(feats_dtype == WFSP_F32 || feats_dtype == WFSP_BF16, "feats dtype must be f32 or bf16");
  cudaStream_t st = as_stream(stream);
  if (n_rows == 0) return WFSP_OK;
  // either half may be skipped (NULL output): the two are independent, so a caller can put them on different streams
  if (indices_bxy != nullptr) {
    WFSP_REQUIRE(coords_xye != nullptr, "indices requested without coordinates");
    pack_indices_kernel<<<unsigned(ceil_div<int64_t>(n_rows, 256)), 256, 0, st>>>(coords_xye, n_rows, n_rows_dev, item_rows,
                                                                                 item_offset, n_items, indices_bxy);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }
  if (feats == nullptr) return WFSP_OK;
  WFSP_REQUIRE(wave != nullptr, "features requested without waveforms");
  if (wave_dtype == WFSP_I16 && feats_dtype == WFSP_F32)
    return launch_pack_feats<int16_t, float>(wave, n_rows, n_rows_dev, n_chan, scale, feats, feats_pitch, st);
  if (wave_dtype == WFSP_I16 && feats_dtype == WFSP_BF16)
    return launch_pack_feats<int16_t, __nv_bfloat16>(wave, n_rows, n_rows_dev, n_chan, scale, feats, feats_pitch, st);
  if (wave_dtype == WFSP_F32 && feats_dtype == WFSP_F32)
    return launch_pack_feats<float, float>(wave, n_rows, n_rows_dev, n_chan, scale, feats, feats_pitch, st);
  return launch_pack_feats<float, __nv_bfloat16>(wave, n_rows, n_rows_dev, n_chan, scale, feats, feats_pitch, st);
}

extern "C" int wfsp_dense_cell_table(const int32_t* indices, int64_t n_rows, const int32_t* n_rows_dev, int batch, int h,
                                     int w, int32_t* cell_table, wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 0 && h > 0 && w > 0 && n_rows >= 0, "bad dense sizes");
  cudaStream_t st = as_stream(stream);
  const int64_t cells = int64_t(batch) * h * w;
  if (cells == 0) return WFSP_OK;
  WFSP_CHECK_CUDA(cudaMemsetAsync(cell_table, 0xff, size_t(cells) * 4, st));
  if (n_rows > 0) {
    dense_mark_kernel<<<unsigned(ceil_div<int64_t>(n_rows, 256)), 256, 0, st>>>(indices, n_rows, n_rows_dev, batch, h, w, cell_table);
    count_launches(1);
    WFSP_CHECK_LAUNCH();
  }
  return WFSP_OK;
}

extern "C" int wfsp_to_dense_from_table(const float* feats, int n_chan, int batch, int h, int w, const int32_t* cell_table,
                                        float* dense, wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 0 && h > 0 && w > 0 && n_chan >= 0, "bad dense sizes");
  const int64_t cells = int64_t(batch) * h * w;
  if (cells == 0 || n_chan == 0) return WFSP_OK;
  dim3 grid(unsigned(ceil_div<int64_t>(cells, 32)), unsigned(ceil_div(n_chan, 32)));
  dense_fwd_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(feats, n_chan, cell_table, cells, h * w, dense);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_to_dense(const float* feats, const int32_t* indices, int64_t n_rows, const int32_t* n_rows_dev,
                             int n_chan, int batch, int h, int w, float* dense, int32_t* cell_table,
                             wfsp_stream_t stream) {
  WFSP_REQUIRE(n_chan >= 0, "bad dense sizes");
  if (n_chan == 0) return WFSP_OK;
  const int rc = wfsp_dense_cell_table(indices, n_rows, n_rows_dev, batch, h, w, cell_table, stream);
  if (rc != WFSP_OK) return rc;
  return wfsp_to_dense_from_table(feats, n_chan, batch, h, w, cell_table, dense, stream);
}

extern "C" int wfsp_to_dense_bwd(const float* d_dense, const int32_t* indices, int64_t n_rows,
                                 const int32_t* n_rows_dev, int n_chan, int batch, int h, int w, float* d_feats,
                                 wfsp_stream_t stream) {
  WFSP_REQUIRE(batch >= 0 && h > 0 && w > 0 && n_chan >= 0 && n_rows >= 0, "bad dense sizes");
  if (n_rows == 0 || n_chan == 0) return WFSP_OK;
  dim3 grid(unsigned(ceil_div<int64_t>(n_rows, 32)), unsigned(ceil_div(n_chan, 32)));
  dense_bwd_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(d_dense, indices, n_rows, n_rows_dev, n_chan, batch, h, w, d_feats);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}


// ---- optimiser step over the flat parameter / gradient buffers ----------------------------------------
// torch.optim.SGD semantics (config/examples/GEP.json:56-68: lr 0.02, momentum 0.98, nesterov):
//   g = grad * grad_scale + weight_decay * p;  buf = momentum * buf + g;
//   p -= lr * (nesterov ? g + momentum * buf : buf)
// with the momentum buffer starting at zero (== torch's "first step copies the gradient", no dampening).
// One streaming launch over all parameters instead of a multi-tensor kernel with a few blocks;
// grad_scale folds the 1 / world_size of the data-parallel mean into the update.
namespace wfsp {
namespace {
__device__ __forceinline__ float sgd_update(float pv, float gv_in, float& b, float lr, float momentum, int nesterov,
                                            float weight_decay, float grad_scale) {
  const float gv = gv_in * grad_scale + weight_decay * pv;
  float step = gv;
  if (momentum != 0.f) {
    b = momentum * b + gv;
    step = nesterov ? gv + momentum * b : b;
  }
  return pv - lr * step;
}

// One 16-byte vector of each buffer per thread and iteration, every load issued before the first store (a store
// between the loads -- e.g. clearing g right after reading it -- holds the later loads back: the scalar version with the
// clearing store took 24 us instead of 5.6).  ZERO: the gradient buffer is cleared in the same pass, so the next step
// accumulates into a clean buffer without a fill launch at its start.  The last n % 4 elements go to one thread.
template <bool ZERO>
__global__ void __launch_bounds__(256) sgd_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                                                       int64_t n, float lr, float momentum, int nesterov, float weight_decay,
                                                       float grad_scale) {
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const bool mom = momentum != 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = p4[i];
    const float4 gv = g4[i];
    float4 bv = mom ? b4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    pv.x = sgd_update(pv.x, gv.x, bv.x, lr, momentum, nesterov, weight_decay, grad_scale);
    pv.y = sgd_update(pv.y, gv.y, bv.y, lr, momentum, nesterov, weight_decay, grad_scale);
    pv.z = sgd_update(pv.z, gv.z, bv.z, lr, momentum, nesterov, weight_decay, grad_scale);
    pv.w = sgd_update(pv.w, gv.w, bv.w, lr, momentum, nesterov, weight_decay, grad_scale);
    p4[i] = pv;
    if (mom) b4[i] = bv;
    if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 << 2; i < n; ++i) {
      float b = mom ? buf[i] : 0.f;
      p[i] = sgd_update(p[i], g[i], b, lr, momentum, nesterov, weight_decay, grad_scale);
      if (mom) buf[i] = b;
      if (ZERO) g[i] = 0.f;
    }
  }
}
struct StageJobs {
  char* dst[3];
  const char* src[3];
  size_t bytes[3];
  int vec[3];  // 16-byte vectors usable (both pointers aligned)
};

// three independent copies in one grid: 16-byte vectors over the aligned body, bytes for the tail
__global__ void __launch_bounds__(256) stage_inputs_kernel(StageJobs j, int32_t* __restrict__ n_dst, int32_t n) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_dst) *n_dst = n;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = size_t(gridDim.x) * blockDim.x;
#pragma unroll 1
  for (int s = 0; s < 3; ++s) {
    if (j.bytes[s] == 0) continue;
    const size_t body = j.vec[s] ? (j.bytes[s] & ~size_t(15)) : 0;
    const uint4* sv = reinterpret_cast<const uint4*>(j.src[s]);
    uint4* dv = reinterpret_cast<uint4*>(j.dst[s]);
#pragma unroll 4
    for (size_t i = tid; i < body / 16; i += nth) dv[i] = sv[i];
#pragma unroll 1
    for (size_t i = body + tid; i < j.bytes[s]; i += nth) j.dst[s][i] = j.src[s][i];
  }
}

}  // namespace
}  // namespace wfsp

extern "C" int wfsp_stage_inputs(void* dst0, const void* src0, size_t bytes0, void* dst1, const void* src1, size_t bytes1,
                                 void* dst2, const void* src2, size_t bytes2, int32_t* n_rows_dev, int32_t n_rows,
                                 wfsp_stream_t stream) {
  wfsp::StageJobs j{};
  void* d[3] = {dst0, dst1, dst2};
  const void* s[3] = {src0, src1, src2};
  const size_t b[3] = {bytes0, bytes1, bytes2};
  size_t total = 0;
  for (int i = 0; i < 3; ++i) {
    const bool on = s[i] != nullptr && b[i] > 0;
    WFSP_REQUIRE(!on || d[i] != nullptr, "staging slot %d has a source but no destination", i);
    j.dst[i] = static_cast<char*>(d[i]);
    j.src[i] = static_cast<const char*>(s[i]);
    j.bytes[i] = on ? b[i] : 0;
    j.vec[i] = on && ((reinterpret_cast<uintptr_t>(d[i]) | reinterpret_cast<uintptr_t>(s[i])) & 15) == 0;
    total += j.bytes[i];
  }
  if (total == 0 && n_rows_dev == nullptr) return WFSP_OK;
  int64_t blocks = wfsp::ceil_div<int64_t>(int64_t(total / 16) + 1, 256 * 4);
  const int64_t cap = int64_t(wfsp::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  wfsp::stage_inputs_kernel<<<unsigned(blocks), 256, 0, wfsp::as_stream(stream)>>>(j, n_rows_dev, n_rows);
  wfsp::count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_sgd_step_ex(float* params, float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                                int nesterov, float weight_decay, float grad_scale, int zero_grads, wfsp_stream_t stream) {
  WFSP_REQUIRE(n >= 0 && params != nullptr && grads != nullptr, "bad optimiser arguments");
  WFSP_REQUIRE(momentum == 0.f || momentum_buf != nullptr, "momentum needs a buffer");
  if (n == 0) return WFSP_OK;
  WFSP_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                 reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0, "flat buffers must be 16-byte aligned");
  int64_t blocks = wfsp::ceil_div<int64_t>(n, 256 * 4);
  const int64_t cap = int64_t(wfsp::sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (zero_grads)
    wfsp::sgd_flat_kernel<true><<<unsigned(blocks), 256, 0, wfsp::as_stream(stream)>>>(params, grads, momentum_buf, n, lr, momentum,
                                                                                       nesterov, weight_decay, grad_scale);
  else
    wfsp::sgd_flat_kernel<false><<<unsigned(blocks), 256, 0, wfsp::as_stream(stream)>>>(params, grads, momentum_buf, n, lr, momentum,
                                                                                        nesterov, weight_decay, grad_scale);
  wfsp::count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

extern "C" int wfsp_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                             int nesterov, float weight_decay, float grad_scale, wfsp_stream_t stream) {
  return wfsp_sgd_step_ex(params, const_cast<float*>(grads), momentum_buf, n, lr, momentum, nesterov, weight_decay, grad_scale,
                          0, stream);
}
