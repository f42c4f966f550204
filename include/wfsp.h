/*
 * wfsp.h -- C ABI of libwfsp.so: the B200 (sm_100a) sparse-convolution hot path of WaveformML.
 *
 * This is the drop-in boundary.  The reference (pure Python) reaches its sparse convolutions
 * through the third-party package spconv~=1.2.1 (/root/reference/requirements.txt:15); the entry
 * points below are what a binding for that path calls instead of upstream's
 * torch.ops.spconv.{get_indice_pairs, indice_conv, indice_conv_backward}.  Each entry cites the
 * reference call site (file:line under /root/reference) whose work it carries out.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *     allocation and no implicit synchronisation; scratch memory is passed in by the caller and
 *     sized with the matching *_workspace_bytes query;
 *   - return value 0 = success, negative = error (WFSP_E*); wfsp_last_error() gives the text of
 *     the calling thread's last failure;
 *   - row-major everywhere; `indices` rows are (batch, x, y) int32 as spconv requires
 *     (src/models/SPConvNet.py:51-52,63-64); features are [rows, channels];
 *   - DEVICE-SIDE ROW COUNTS: every row count `n_x` has a companion `const int32_t* n_x_dev`.
 *     NULL: `n_x` is the exact count (the eager, reference-shaped path).  Non-NULL: `n_x` is only
 *     the CAPACITY of the buffers / the launch bound and the kernels read the live count from
 *     *n_x_dev, so a whole training step can be enqueued without any host readback and replayed
 *     from a CUDA graph; rows past the live count are neither read nor written;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns WFSP_ECUDA.
 */
#ifndef WFSP_H_
#define WFSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* wfsp_stream_t; /* cudaStream_t */

enum wfsp_status {
  WFSP_OK = 0,
  WFSP_EINVAL = -1,       /* bad argument */
  WFSP_ECUDA = -2,        /* CUDA runtime error (text in wfsp_last_error) */
  WFSP_EWORKSPACE = -3,   /* workspace too small */
  WFSP_EUNSUPPORTED = -4  /* shape outside what the kernels cover */
};

enum wfsp_dtype { WFSP_F32 = 0, WFSP_BF16 = 1, WFSP_I16 = 2 };

/* arithmetic of the channel contraction */
enum wfsp_math {
  WFSP_MATH_FP32 = 0, /* fp32 FMA on CUDA cores, exact fp32 products                      */
  WFSP_MATH_BF16 = 1, /* bf16 operands, fp32 accumulate in TMEM, tcgen05.mma (kind::f16)  */
  WFSP_MATH_BF16X3 = 2 /* fp32-grade results on the same tensor-core kernels: every operand is split into
                          hi + lo bf16 parts and a w ~ hi hi + hi lo + lo hi is accumulated in fp32 (three
                          times the reduction length; error ~2^-16 per product, vs 2^-9 for WFSP_MATH_BF16).
                          The tensor-core counterpart of WFSP_MATH_FP32 for the tight-tolerance mode.    */
};

#define WFSP_VERSION 203
#define WFSP_MAX_KVOL 1024

int wfsp_version(void);
const char* wfsp_last_error(void);
/* sm count and compute capability of the current device */
int wfsp_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);
/* tuning / test knobs; "rulebook_force_hash" = 1 forces the open-addressing coordinate hash even
 * when the dense grid table would be used */
int wfsp_set_option(const char* name, int value);
/* Debugging aid: while device_buffer (>= 128 + 2048 uint64) is set, tile (0, 0) of every wfsp_conv_apply_bf16[_ex]
 * launch stamps clock64() at its phase boundaries (slot 15 of each cluster rank: %globaltimer at entry), and every live
 * CTA of the first 1024 writes %globaltimer at entry / exit to slots 128 + 2 * linear block id (+1).  NULL switches it
 * off (the default; costs one predicate per phase). */
int wfsp_debug_trace(unsigned long long* device_buffer);
/* number of CUDA kernels this library has launched in the calling process (monotonic) */
unsigned long long wfsp_kernel_launches(void);
/* hex digest of the sources (csrc/ + this header) the library was compiled from; the Python loader compares it
 * with the tree it runs in and rebuilds (or refuses) when they differ -- a stale binary never gets tested */
const char* wfsp_source_hash(void);

/* ---------------------------------------------------------------------------------------------
 * (1) Sparse-tensor batcher.
 * Replaces collate_fn (src/engineering/PSDDataModule.py:10-20: add the running event offset to
 * coords[:,2] of every item after the first, concatenate), the dtype / normalisation part of
 * HDF5Dataset._concat_range (src/datasets/HDF5Dataset.py:282-302 coords->int32, :227,290
 * waveform->float, :345-346 `vals *= 1/(2^14-1)`), and the batch-first column permute of
 * SPConvNet.forward (src/models/SPConvNet.py:63-64, `x[0][:, [2,0,1]]`).
 *
 *   coords_xye   int32 [n_rows,3] = (x, y, event id local to its item), items concatenated
 *   wave         [n_rows, n_chan] of wave_dtype (WFSP_I16 as stored on disk,
 *                src/datasets/H5CompoundTypes.py:105-120, or WFSP_F32)
 *   item_rows    int64 [n_items+1] row offset of each item
 *   item_offset  int64 [n_items]   event offset added to each item (0 for item 0)
 *   indices_bxy  int32 [n_rows,3] = (global event, x, y)                       (output)
 *   feats        [n_rows, feats_pitch] of feats_dtype, value = wave * scale    (output)
 *   Either output may be NULL (with its input): the two halves are independent, so a caller may issue them
 *   on different streams (the rulebooks only wait for the indices).
 */
int wfsp_batch_pack(const int32_t* coords_xye, const void* wave, int wave_dtype, int64_t n_rows,
                    const int32_t* n_rows_dev, int n_chan, const int64_t* item_rows,
                    const int64_t* item_offset, int64_t n_items, float scale, int32_t* indices_bxy,
                    void* feats, int feats_dtype, int64_t feats_pitch, wfsp_stream_t stream);

/* Staging of one batch into the capacity-sized input buffers of a captured step (the graph path's
 * counterpart of `.to(device)` in HDF5Dataset._concat_range, src/datasets/HDF5Dataset.py:340-346):
 * copies up to three device buffers (coords, waveforms, labels; a NULL source skips its slot) and
 * writes the live row count, in ONE launch -- at 64 events four separate copies cost more than any
 * kernel of the step.  All pointers are device pointers; byte counts need no alignment. */
int wfsp_stage_inputs(void* dst0, const void* src0, size_t bytes0, void* dst1, const void* src1,
                      size_t bytes1, void* dst2, const void* src2, size_t bytes2, int32_t* n_rows_dev,
                      int32_t n_rows, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (2) Rulebook builder.  Bit-exact with the CPU path of upstream ops.get_indice_pairs, which
 * every spconv.SparseConv2d / SubMConv2d with kernel volume > 1 reaches
 * (src/models/SPConvBlocks.py:75,134,191,249,298,370,498-502,575-579,643-647,710-714,803,809,
 * 868,877,930,939): output rows in first-touch order, pairs of one offset in ascending input
 * order, pairs tensor [2, kvol, n_in] padded with -1, pair_num [kvol].
 */
/* out = (in + 2p - d(k-1) - 1)/s + 1 per dim (same as src/utils/ModelValidation.py:119-126) */
int wfsp_conv_out_shape(const int* in_shape_host, const int* ksize_host, const int* stride_host,
                        const int* pad_host, const int* dil_host, int* out_shape_host);

size_t wfsp_rulebook_workspace_bytes(int64_t n_in, int batch, const int* out_shape_host,
                                     const int* ksize_host);

/* Regular (strided / padded / dilated) convolution.  out_indices must hold out_cap >=
 * min(n_in*kvol, batch*out_h*out_w) rows; the number of rows actually produced is written to the
 * device scalar *n_out (in the eager path the caller copies it to the host -- the single readback
 * per rulebook; in the graph path it feeds the next call's n_*_dev).  stride>1 together with
 * dilation>1 is rejected as upstream does.  The pair arrays always have pitch n_in. */
int wfsp_rulebook_conv(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                       const int* in_shape_host, const int* ksize_host, const int* stride_host,
                       const int* pad_host, const int* dil_host, int32_t* out_indices,
                       int64_t out_cap, int32_t* pairs, int32_t* pair_num, int32_t* n_out,
                       void* workspace, size_t workspace_bytes, wfsp_stream_t stream);

/* Submanifold convolution: output rows == input rows, padding forced to k/2, stride to 1. */
int wfsp_rulebook_subm(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                       const int* shape_host, const int* ksize_host, const int* dil_host,
                       int32_t* pairs, int32_t* pair_num, void* workspace, size_t workspace_bytes,
                       wfsp_stream_t stream);

/* Derived, output-stationary view of a rulebook used by the forward / dgrad kernels:
 *   nbr_out [n_out, kvol]: input row that feeds output row o through offset k, or -1
 *   nbr_in  [n_in,  kvol]: output row that input row i feeds through offset k, or -1
 * n_in / n_out may be capacities (only pair_num[k] pairs per offset are read).
 * *dup_flag (device int32, caller zero-fills) is set to 1 if two pairs of one offset share an
 * output row (duplicate input coordinates), which the output-stationary kernels do not cover. */
int wfsp_rulebook_tables(const int32_t* pairs, const int32_t* pair_num, int kvol,
                         int64_t pair_pitch, int64_t n_in, int64_t n_out, int32_t* nbr_out,
                         int32_t* nbr_in, int32_t* dup_flag, wfsp_stream_t stream);

/* Rulebook and neighbour tables in one call: wfsp_rulebook_conv (subm == 0) or wfsp_rulebook_subm
 * (subm != 0; stride / pad / out_indices / n_out ignored, may be NULL) followed by
 * wfsp_rulebook_tables, with nbr_out sized [out_cap, kvol] (subm: [n_in, kvol]).  *dup_flag is
 * written (0 / 1), no need to clear it.  Inputs of up to 2048 (expected live) rows are built by a
 * single kernel launch (no memsets): at the reference's batch size of 64 events the step is
 * launch-latency bound.  n_in_hint (graph path; 0 = none): expected live rows, launch shaping only. */
int wfsp_rulebook_build(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev,
                        int64_t n_in_hint, int batch,
                        const int* in_shape_host, const int* ksize_host, const int* stride_host,
                        const int* pad_host, const int* dil_host, int subm, int32_t* out_indices,
                        int64_t out_cap, int32_t* pairs, int32_t* pair_num, int32_t* n_out,
                        int32_t* nbr_out, int32_t* nbr_in, int32_t* dup_flag, void* workspace,
                        size_t workspace_bytes, wfsp_stream_t stream);

/* The same three calls for ndim = 2 or 3 spatial dimensions.  ndim 3 is the reference's net_type
 * "3DConvolution" (src/models/SPConvNet.py:42-49, src/models/SCNet.py:53-55: spatial size
 * [14, 11, n_samples], index rows (b, x, y, t) = int32 [n, 4]); all *_host arrays hold ndim ints, kernel
 * offsets are numbered row-major over (kx, ky, kt) as upstream does, out_indices has ndim + 1 columns.
 * The rulebook is what spconv.SparseConv3d / SubMConv3d (src/utils/ModelValidation.py:24-31) would build;
 * wfsp_conv_apply* / wfsp_conv_wgrad* consume it unchanged (they only see kvol and the tables). */
int wfsp_conv_out_shape_nd(int ndim, const int* in_shape_host, const int* ksize_host,
                           const int* stride_host, const int* pad_host, const int* dil_host,
                           int* out_shape_host);
size_t wfsp_rulebook_workspace_bytes_nd(int ndim, int64_t n_in, int batch, const int* out_shape_host,
                                        const int* ksize_host);
int wfsp_rulebook_build_nd(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev,
                           int64_t n_in_hint, int batch, const int* in_shape_host,
                           const int* ksize_host, const int* stride_host, const int* pad_host,
                           const int* dil_host, int subm, int32_t* out_indices, int64_t out_cap,
                           int32_t* pairs, int32_t* pair_num, int32_t* n_out, int32_t* nbr_out,
                           int32_t* nbr_in, int32_t* dup_flag, void* workspace, size_t workspace_bytes,
                           wfsp_stream_t stream);

/* wfsp_rulebook_build_nd in two launches, for callers that overlap the backward half of a rulebook with the forward
 * pass (waveformml_b200/spconv/fused.py): phases = 1 (FRONT) produces out_indices, *n_out and nbr_out -- all the forward
 * convolution and the next layer's rulebook wait for; phases = 2 (BACK, same arguments, any stream ordered after the
 * front call) produces pairs, pair_num, nbr_in and dup_flag; phases = 3 = everything (= wfsp_rulebook_build_nd).  Only
 * the single-launch builder (small inputs) is split: otherwise the FRONT call builds everything and sets *built_all
 * (host int, may be NULL) and the BACK call returns immediately. */
int wfsp_rulebook_build_phased(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev,
                               int64_t n_in_hint, int batch, const int* in_shape_host,
                               const int* ksize_host, const int* stride_host, const int* pad_host,
                               const int* dil_host, int subm, int32_t* out_indices, int64_t out_cap,
                               int32_t* pairs, int32_t* pair_num, int32_t* n_out, int32_t* nbr_out,
                               int32_t* nbr_in, int32_t* dup_flag, void* workspace, size_t workspace_bytes,
                               int phases, int* built_all, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Gather-GEMM forward / dgrad / wgrad.  Replaces upstream indice_conv /
 * indice_conv_backward reached from spconv.SparseConv2d / SubMConv2d / SparseInverseConv2d
 * .forward and their autograd backward (same call sites as (2), plus the 1x1 shortcut
 * `torch.mm(features, weight.view(Cin,Cout))` of every pointwise layer,
 * src/models/SPConvBlocks.py:335,498).
 *
 * wfsp_conv_apply computes, for every destination row r,
 *     dst[r, :] = bias + sum_k  src[ nbr[r,k], : ] @ Wk          (rows with nbr = -1 contribute 0)
 * where Wk = weight[k] ([c_red, c_dst]) if transpose_w == 0, or weight[k]^T if transpose_w == 1
 * (weight[k] is then [c_dst, c_red]).  The four uses:
 *     forward            nbr = nbr_out, src = features, weight [kvol,Cin,Cout], transpose_w = 0
 *     dgrad              nbr = nbr_in,  src = dOut,     weight [kvol,Cin,Cout], transpose_w = 1
 *     inverse forward    nbr = nbr_in,  src = features, weight [kvol,Cin,Cout], transpose_w = 0
 *     inverse dgrad      nbr = nbr_out, src = dOut,     weight [kvol,Cin,Cout], transpose_w = 1
 * nbr == NULL means kvol == 1 with the identity map (the 1x1 shortcut, n_dst == n_src).
 * src and dst are fp32 [rows, channels] with row pitch == channels; weight and bias are fp32.
 */
size_t wfsp_conv_apply_workspace_bytes(int kvol, int64_t n_src, int c_red, int c_dst, int math);

/* n_dst_hint (graph path; 0 = none): the number of live destination rows the caller expects, used
 * only to pick the launch shape (column tiling) -- never for correctness. */
int wfsp_conv_apply(const float* src, int64_t n_src, const int32_t* n_src_dev, int c_red,
                    const float* weight, int transpose_w, const float* bias, const int32_t* nbr,
                    int kvol, float* dst, int64_t n_dst, const int32_t* n_dst_dev,
                    int64_t n_dst_hint, int c_dst, int math, void* workspace,
                    size_t workspace_bytes, wfsp_stream_t stream);

/* wgrad: d_weight[k] (+)= sum over pairs p of offset k of  a[pa[k,p], :]^T (outer) b[pb[k,p], :]
 * with a = features [n_a, c_a], b = dOut [n_b, c_b], d_weight fp32 [kvol, c_a, c_b]
 * (pair_a = pairs[0], pair_b = pairs[1]; for the inverse convolution the roles swap:
 * pair_a = pairs[1], pair_b = pairs[0]).  pair_a / pair_b are [kvol, pair_pitch], only the first
 * pair_num[k] entries of row k are read.  d_weight is overwritten (accumulate == 0) or added to
 * (accumulate == 1).  pair_a == pair_b == pair_num == NULL selects the identity pair list of the
 * 1x1 shortcut (kvol == 1, n_a == n_b pairs). */
size_t wfsp_conv_wgrad_workspace_bytes(int kvol, int64_t n_a, int c_a, int64_t n_b, int c_b,
                                       int64_t pair_pitch, int math);

/* pairs_hint (graph path; 0 = none): expected pairs of the fullest offset, launch shaping only. */
int wfsp_conv_wgrad(const float* a, int64_t n_a, const int32_t* n_a_dev, int c_a, const float* b,
                    int64_t n_b, const int32_t* n_b_dev, int c_b, const int32_t* pair_a,
                    const int32_t* pair_b, const int32_t* pair_num, int kvol, int64_t pair_pitch,
                    int64_t pairs_hint, float* d_weight, int accumulate, int math, void* workspace,
                    size_t workspace_bytes, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (4) Dense scatter.  Replaces spconv.ToDense / SparseConvTensor.dense()
 * (src/models/SPConvBlocks.py:81,726; src/engineering/LitBase.py:138-146): zeros
 * [B, C, H, W], dense[b, :, x, y] = features[row] (last row wins on duplicate coordinates), and
 * its backward, the gather d_features[row, :] = d_dense[b, :, x, y].
 * cell_table: int32 scratch [batch*h*w].
 */
int wfsp_to_dense(const float* feats, const int32_t* indices, int64_t n_rows,
                  const int32_t* n_rows_dev, int n_chan, int batch, int h, int w, float* dense,
                  int32_t* cell_table, wfsp_stream_t stream);

/* The two halves of wfsp_to_dense.  The cell table (row of every dense cell, -1 = empty) depends on the
 * coordinates only, so a caller that knows the output coordinates early (the rulebook of the last layer)
 * builds it beside the convolutions; the scatter then is one launch behind the last layer. */
int wfsp_dense_cell_table(const int32_t* indices, int64_t n_rows, const int32_t* n_rows_dev, int batch,
                          int h, int w, int32_t* cell_table, wfsp_stream_t stream);
int wfsp_to_dense_from_table(const float* feats, int n_chan, int batch, int h, int w,
                             const int32_t* cell_table, float* dense, wfsp_stream_t stream);

int wfsp_to_dense_bwd(const float* d_dense, const int32_t* indices, int64_t n_rows,
                      const int32_t* n_rows_dev, int n_chan, int batch, int h, int w,
                      float* d_feats, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (5) BatchNorm1d (+ReLU) over the active rows -- the modules the reference's SparseSequential
 * applies to `.features` between convolutions (src/models/SPConvBlocks.py:505-508:
 * nn.BatchNorm1d(c), nn.ReLU()).  Needed natively only by the graph path, where the row count
 * lives on the device; same arithmetic as torch.nn.BatchNorm1d in training mode (biased batch
 * variance for normalisation, unbiased for the running estimate).
 *   forward : y = relu?((x - mean) * invstd * gamma + beta); saves mean / invstd [c];
 *             running_mean / running_var (may be NULL) updated with `momentum`.
 *             training == 0 normalises with the running statistics instead.
 *   backward: dx, d_gamma, d_beta from dy (ReLU mask recomputed from x).
 * workspace: wfsp_bn_workspace_bytes(n_rows, c) bytes.
 */
size_t wfsp_bn_workspace_bytes(int64_t n_rows, int c);

int wfsp_bn_relu_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c,
                     const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int training, int relu,
                     float* y, float* save_mean, float* save_invstd, void* workspace,
                     size_t workspace_bytes, wfsp_stream_t stream);

int wfsp_bn_relu_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev,
                     int c, const float* gamma, const float* beta, const float* save_mean,
                     const float* save_invstd, int relu, float* dx, float* d_gamma, float* d_beta,
                     void* workspace, size_t workspace_bytes, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (6) bf16-resident pipeline.  The entries of (3) and (5) take and return fp32 [rows, channels]
 * tensors, which is what the torch modules between the reference's sparse layers exchange
 * (src/models/SPConvBlocks.py:498-508: SparseConv2d, nn.BatchNorm1d, nn.ReLU).  A caller that owns a
 * whole SparseSequential stack (waveformml_b200/spconv/fused.py) can instead keep the tensor-core
 * operands resident in bf16 between layers: every activation is cast ONCE, by the kernel that
 * produces it, the weights of all layers are prepared by one launch per step, and forward /
 * dgrad / wgrad read those copies directly.  Same arithmetic as WFSP_MATH_BF16 in (3): bf16 operands,
 * fp32 accumulation, fp32 results.
 *
 * bf16 activation format: [rows, WFSP_BF16_PITCH(c)] row-major, channels c .. pitch-1 zero.
 * Prepared weight format: opaque, wfsp_prepared_weight_bytes() bytes, made by wfsp_prep_weights for
 * one (weight, direction): transpose_w = 0 for forward / inverse forward, 1 for dgrad.
 */
#define WFSP_BF16_PITCH(c) (((c) + 7) / 8 * 8)

typedef struct wfsp_prep_job {
  const float* weight; /* fp32 [kvol, c_red, c_dst], or [kvol, c_dst, c_red] if transpose_w      */
  void* out;           /* prepared bf16 weights, wfsp_prepared_weight_bytes(kvol, c_red, c_dst); 16-byte aligned */
  int kvol, c_red, c_dst, transpose_w;
} wfsp_prep_job;

size_t wfsp_prepared_weight_bytes(int kvol, int c_red, int c_dst);
/* jobs_host: HOST array; one kernel launch per 16 jobs (sized to leave about a third of the SMs free: it is meant
 * to run beside the first convolution of a step, whose CTAs each need a whole SM) */
int wfsp_prep_weights(const wfsp_prep_job* jobs_host, int n_jobs, wfsp_stream_t stream);

/* fp32 [n_rows, c] -> bf16 [n_rows, WFSP_BF16_PITCH(c)] */
int wfsp_cast_rows_bf16(const float* src, int64_t n_rows, const int32_t* n_rows_dev, int c,
                        void* dst_bf16, wfsp_stream_t stream);

/* wfsp_conv_apply on bf16 activations and prepared weights (no workspace, no cast pass).
 * bn_partials (may be NULL): wfsp_bn_partials_bytes(n_dst, c_dst) bytes, starting with fp32
 * [ceil(n_dst / 32)][2][c_dst]; the epilogue writes (mean, M2) of every
 * output column over each chunk of 32 destination rows, which wfsp_bn_relu_fwd_stats turns into the
 * BatchNorm statistics without re-reading dst (WFSP_BN_CHUNK_ROWS). */
#define WFSP_BN_CHUNK_ROWS 32
int wfsp_conv_apply_bf16(const void* src_bf16, int64_t n_src, const int32_t* n_src_dev, int c_red,
                         const void* weight_prepared, const float* bias, const int32_t* nbr, int kvol,
                         float* dst, int64_t n_dst, const int32_t* n_dst_dev, int64_t n_dst_hint,
                         int c_dst, float* bn_partials, wfsp_stream_t stream);

/* wfsp_conv_apply_bf16 with the optional fused epilogues / launch options of the bf16-resident stack (fused.py);
 * a zeroed struct (or NULL) is plain wfsp_conv_apply_bf16 without bn_partials.
 *   bn_partials   as above (forward: statistics of the BatchNorm that FOLLOWS this convolution,
 *                 src/models/SPConvBlocks.py:505-508).
 *   bwd_*         dgrad only: the output of this call is the gradient dy arriving at a BatchNorm(+ReLU) whose
 *                 input was bwd_x [n_dst, c_dst] with saved statistics bwd_mean / bwd_invstd and affine parameters
 *                 bwd_gamma / bwd_beta (may be NULL).  The epilogue then also writes, per chunk of 32 destination
 *                 rows, bwd_partials [ceil(n_dst / 32)][2][c_dst] = (sum dy', sum dy' * xhat) with dy' = dy masked
 *                 by the ReLU -- the two reductions of BatchNorm backward, taken from the tile while it is on
 *                 chip, so wfsp_bn_relu_bwd_parts needs no reduction pass over (x, dy).  Buffer size:
 *                 wfsp_bn_partials_bytes(n_dst, c_dst).
 *   k_split       0 = automatic.  Small launches (a handful of tiles) are bound by the serial walk of each CTA over
 *                 (kernel offset x 64-channel slice); there a thread-block cluster of up to 8 CTAs splits that
 *                 loop and the partial accumulators are added through distributed shared memory in rank order
 *                 (deterministic).  1 = never split; 2 / 4 / 8 = force (tests). */
struct wfsp_dropout;
/* The BatchNorm1d(+ReLU, +Dropout) that FOLLOWS a convolution, finished inside the convolution's launch (small
 * launches: all CTAs co-resident; a grid-wide barrier separates the statistics from the normalisation).  When the
 * launch cannot take it (too many CTAs, large input) the library runs wfsp_bn_relu_fwd_stats_ex right behind the
 * convolution instead: the caller gets the same results either way and never launches the BatchNorm itself.
 * Fields as in wfsp_bn_relu_fwd_stats_ex; barrier: device uint32 [2], zero before the first use (reuse it). */
typedef struct wfsp_bn_fuse {
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float momentum, eps;
  int relu;
  float* y;
  void* y_bf16;
  float* save_mean;
  float* save_invstd;
  const struct wfsp_dropout* dropout;
  void* barrier;
  int64_t n_rows_hint;
} wfsp_bn_fuse;

typedef struct wfsp_conv_epilogue {
  float* bn_partials;
  const float* bwd_x;
  const float* bwd_mean;
  const float* bwd_invstd;
  const float* bwd_gamma;
  const float* bwd_beta;
  float* bwd_partials;
  int bwd_relu;
  int k_split;
  const wfsp_bn_fuse* bn; /* needs bn_partials */
} wfsp_conv_epilogue;
int wfsp_conv_apply_bf16_ex(const void* src_bf16, int64_t n_src, const int32_t* n_src_dev, int c_red,
                            const void* weight_prepared, const float* bias, const int32_t* nbr, int kvol,
                            float* dst, int64_t n_dst, const int32_t* n_dst_dev, int64_t n_dst_hint,
                            int c_dst, const wfsp_conv_epilogue* epilogue, wfsp_stream_t stream);

/* wfsp_conv_wgrad on bf16 rows (a = layer input, b = gradient of the layer output) */
int wfsp_conv_wgrad_bf16(const void* a_bf16, int64_t n_a, const int32_t* n_a_dev, int c_a,
                         const void* b_bf16, int64_t n_b, const int32_t* n_b_dev, int c_b,
                         const int32_t* pair_a, const int32_t* pair_b, const int32_t* pair_num,
                         int kvol, int64_t pair_pitch, int64_t pairs_hint, float* d_weight,
                         int accumulate, wfsp_stream_t stream);

/* BatchNorm1d(+ReLU) as in (5) with optional outputs: y / dx fp32 [rows, c] and / or y_bf16 / dx_bf16
 * in the bf16 activation format (either may be NULL, not both). */
int wfsp_bn_relu_fwd_x(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c,
                       const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int training, int relu,
                       float* y, void* y_bf16, float* save_mean, float* save_invstd,
                       void* workspace, size_t workspace_bytes, wfsp_stream_t stream);

/* size of the bn_partials buffer of wfsp_conv_apply_bf16 / wfsp_bn_relu_fwd_stats for n_rows output
 * rows (the partial list followed by the scratch of the merge kernel) */
size_t wfsp_bn_partials_bytes(int64_t n_rows, int c);

/* training-mode forward whose per-chunk statistics were already written by wfsp_conv_apply_bf16
 * (bn_partials, chunks of WFSP_BN_CHUNK_ROWS rows): merge + normalise, no statistics pass over x.
 * n_rows_hint (here and in wfsp_bn_relu_bwd_x; graph path, 0 = none): expected live rows -- only
 * selects between the single-launch variants for few rows and the streaming kernels. */
int wfsp_bn_relu_fwd_stats(const float* x, int64_t n_rows, const int32_t* n_rows_dev,
                           int64_t n_rows_hint, int c,
                           const float* bn_partials, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps,
                           int relu, float* y, void* y_bf16, float* save_mean, float* save_invstd,
                           wfsp_stream_t stream);

int wfsp_bn_relu_bwd_x(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev,
                       int64_t n_rows_hint, int c, const float* gamma, const float* beta, const float* save_mean,
                       const float* save_invstd, int relu, float* dx, void* dx_bf16, float* d_gamma,
                       float* d_beta, void* workspace, size_t workspace_bytes, wfsp_stream_t stream);

/* BatchNorm1d(+ReLU) backward whose two reductions were already taken by the dgrad epilogue
 * (wfsp_conv_epilogue.bwd_partials, chunks of WFSP_BN_CHUNK_ROWS rows): fold the partials (fixed order, double
 * precision) into d_gamma / d_beta and stream dx / dx_bf16 -- one pass over (x, dy) instead of two. */
int wfsp_bn_relu_bwd_parts(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev,
                           int64_t n_rows_hint, int c, const float* gamma, const float* beta,
                           const float* save_mean, const float* save_invstd, int relu,
                           const float* bwd_partials, float* dx, void* dx_bf16, float* d_gamma,
                           float* d_beta, wfsp_stream_t stream);

/* Training-mode nn.Dropout(p) behind a block's BatchNorm . ReLU (src/models/SPConvBlocks.py:375-376, 509-510), fused
 * into the kernels that produce / consume the block's output.  Element (row, channel) is kept with probability 1 - p
 * and scaled by 1 / (1 - p) (torch.nn.Dropout semantics); which elements are kept is a counter-based hash of
 * (seed, *step_dev, salt, row * c + channel) -- the backward pass regenerates the mask, nothing is stored.  step_dev
 * (device int64, may be NULL) lets a captured CUDA graph draw a fresh mask every replay: the caller advances it once
 * per step.  The random stream is NOT torch's Philox stream (parity is distributional, as between any two seeds).
 * wfsp_dropout_factors writes the factor (0 or 1 / (1 - p)) of every element of an [n_rows, c] tensor (tests). */
typedef struct wfsp_dropout {
  float p;
  unsigned long long seed;
  const void* step_dev;
  unsigned salt;
} wfsp_dropout;
int wfsp_dropout_factors(const wfsp_dropout* drop, int64_t n_rows, int c, float* factors, wfsp_stream_t stream);
/* wfsp_bn_relu_fwd_stats / wfsp_bn_relu_bwd_x with Dropout behind the ReLU (drop == NULL or p == 0: none).  Forward:
 * y = dropout(relu?(bn(x))); backward: dy is first multiplied by the regenerated mask. */
int wfsp_bn_relu_fwd_stats_ex(const float* x, int64_t n_rows, const int32_t* n_rows_dev,
                              int64_t n_rows_hint, int c,
                              const float* bn_partials, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float momentum, float eps,
                              int relu, float* y, void* y_bf16, float* save_mean, float* save_invstd,
                              const wfsp_dropout* drop, wfsp_stream_t stream);
int wfsp_bn_relu_bwd_x_ex(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev,
                          int64_t n_rows_hint, int c, const float* gamma, const float* beta, const float* save_mean,
                          const float* save_invstd, int relu, float* dx, void* dx_bf16, float* d_gamma,
                          float* d_beta, void* workspace, size_t workspace_bytes, const wfsp_dropout* drop,
                          wfsp_stream_t stream);

/* a convolution followed by nn.ReLU (or nothing) without BatchNorm: y = relu?(x), and its backward
 * dx = dy * (x > 0 or no relu), with the same optional outputs */
int wfsp_act_fwd(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, int relu,
                 float* y, void* y_bf16, wfsp_stream_t stream);
int wfsp_act_bwd(const float* x, const float* dy, int64_t n_rows, const int32_t* n_rows_dev, int c,
                 int relu, float* dx, void* dx_bf16, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (7) Optimiser step over flat buffers.  The reference trains with torch.optim.SGD (lr 0.02, momentum
 * 0.98, nesterov: config/examples/GEP.json:56-68) after Lightning's DDP gradient mean
 * (src/utils/util.py:233-236).  With all parameters / gradients in one flat fp32 buffer each
 * (harness.FlatGrads) the update is one streaming launch:
 *   g = grads * grad_scale + weight_decay * p;  buf = momentum * buf + g;
 *   p -= lr * (nesterov ? g + momentum * buf : buf)         (buf starts at zero; no dampening)
 * grad_scale = 1 / world_size folds the data-parallel mean into the update. */
int wfsp_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr,
                  float momentum, int nesterov, float weight_decay, float grad_scale,
                  wfsp_stream_t stream);
/* zero_grads != 0: the gradient buffer is cleared in the same pass (optimizer.zero_grad() folded in), so a training
 * step that accumulates into it needs no fill launch at its start. */
int wfsp_sgd_step_ex(float* params, float* grads, float* momentum_buf, int64_t n, float lr,
                     float momentum, int nesterov, float weight_decay, float grad_scale,
                     int zero_grads, wfsp_stream_t stream);

/* The same update for data-parallel training (one process per GPU), fused with the gradient exchange over NVLink peer
 * memory -- replaces NCCL all-reduce + optimiser step (src/utils/util.py:233-236: Lightning DDP).  The flat gradient
 * and parameter buffers of every rank are peer-mapped; rank r sums shard r of all ranks' gradients in rank order
 * (reduce-scatter), updates shard r of the parameters (momentum kept by rank r only), and stores the new values into
 * every rank's parameter buffer (all-gather): ONE launch plus a one-warp launch that completes the closing barrier.
 * All ranks end with bit-identical parameters.
 *   peer_grads_dev / peer_params_dev / peer_flags_dev   DEVICE arrays of `world` pointers (this rank included) to the
 *       peers' flat gradient buffers, flat parameter buffers and flag arrays (uint32 [2 * world], zero before the
 *       first call on every rank)
 *   flags    this rank's own flag array; state: LOCAL uint32 [4], zero before the first call (step epoch, CTA ticket, barrier pending)
 *   grad_scale = 1 / world folds the data-parallel mean in; every rank must call this once per step.
 *   wait_now = 1: the closing barrier is completed here (parameters final and gradients reusable when the stream
 *       reaches the end of the call).  wait_now = 0: the caller completes it with wfsp_sgd_p2p_wait before anything
 *       reads the parameters or overwrites the gradients -- e.g. at the START of the next step, where the other ranks'
 *       stragglers are hidden behind that step's input handling. */
int wfsp_sgd_step_p2p(float* params, const float* grads, float* momentum_buf, int64_t n, float lr,
                      float momentum, int nesterov, float weight_decay, float grad_scale,
                      const void* peer_grads_dev, const void* peer_params_dev,
                      const void* peer_flags_dev, void* flags, void* state, int rank, int world,
                      int wait_now, wfsp_stream_t stream);
int wfsp_sgd_p2p_wait(void* flags, void* state, int world, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (8) Dense classification head + loss.  SPConvNet flattens the ToDense output and applies
 * LinearBlock = Linear(k0, h1) . Linear(h1, n_class), no activation in between
 * (src/models/SPConvNet.py:67-68, src/models/ConvBlocks.py:82-102); LitPSD.training_step takes
 * CrossEntropyLoss (mean) of it (src/engineering/LitPSD.py:94-104).  At 64 events those ~25 library
 * kernels are >40 % of a training step; here they are three launches, same fp32 arithmetic.
 *   wfsp_head_ce_fwd: h1 = x w1^T + b1; logits = h1 w2^T + b2; *loss = mean CE(logits, labels); and the
 *     small half of the backward pass for d loss = 1: dlogits [B,C], dh1 [B,h1], dw2 [C,h1], db2 [C].
 *   wfsp_head_bwd: dx = dh1 w1 [B,k0] (dx may be NULL), dw1 = dh1^T x [h1,k0], db1 (may be NULL), and
 *     dw2 / db2 copied from the forward's values -- all multiplied by *grad_out (device scalar, NULL = 1).
 * Limits: h1 <= 128, n_class <= 64; one CTA reduces the batch, so intended for batch <= ~256.
 * x [B,k0], w1 [h1,k0], w2 [n_class,h1] row-major fp32 (torch.nn.Linear layout); labels int64. */
size_t wfsp_head_workspace_bytes(int batch, int k0, int h1);
int wfsp_head_ce_fwd(const float* x, const float* w1, const float* b1, const float* w2,
                     const float* b2, const int64_t* labels, int batch, int k0, int h1_dim,
                     int n_class, float* h1, float* logits, float* loss, float* dlogits, float* dh1,
                     float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                     wfsp_stream_t stream);
/* The tail of wfsp_head_ce_fwd for ANY batch size, for callers that computed h1 = x w1^T + b1 themselves (one plain GEMM:
 * library): logits, mean cross-entropy, dlogits, dh1, dw2, db2 in ONE launch of ceil(batch / 16) CTAs; the batch
 * reductions (loss, dw2, db2) are added in tile order by the last CTA to finish (deterministic).  *ticket: a device
 * counter that is zero before the call and is left zero (allocate once, reuse).  Workspace:
 * wfsp_head_tail_workspace_bytes(). */
size_t wfsp_head_tail_workspace_bytes(int batch, int h1, int n_class);
int wfsp_head_ce_tail(const float* h1, const float* w2, const float* b2, const int64_t* labels, int batch,
                      int h1_dim, int n_class, float* logits, float* loss, float* dlogits, float* dh1,
                      float* dw2, float* db2, void* workspace, size_t workspace_bytes, unsigned* ticket,
                      wfsp_stream_t stream);
/* backward split for callers with a GEMM library at hand: this entry does the small part (dh1_scaled = dh1 *
 * *grad_out, db1, dw2, db2); the caller then computes dw1 = dh1_scaled^T x and dx = dh1_scaled w1 as two
 * plain GEMMs (what waveformml_b200/head.py does with cuBLAS -- the all-in-one wfsp_head_bwd below is an
 * fp32 SIMT kernel that the library GEMMs beat). */
int wfsp_head_bwd_small(const float* dh1, const float* dw2_in, const float* db2_in,
                        const float* grad_out, int batch, int h1_dim, int n_class, float* dh1_scaled,
                        float* db1, float* dw2, float* db2, wfsp_stream_t stream);
int wfsp_head_bwd(const float* x, const float* w1, const float* dh1, const float* dw2_in,
                  const float* db2_in, const float* grad_out, int batch, int k0, int h1_dim,
                  int n_class, float* dx, float* dw1, float* db1, float* dw2, float* db2,
                  wfsp_stream_t stream);

/* Masked L1 segment loss of the z / energy regression models (src/engineering/LitBase.py:124-174
 * _calc_segment_loss; src/engineering/LitZ.py:89-107): the reference densifies a ones-mask and the per-hit target with
 * SparseConvTensor.dense() and takes l1_loss(mask * prediction, target, "sum") / N.  Inactive cells contribute
 * |0 - 0|, so the same value is the row-wise L1 between the dense prediction at every hit's cell and the hit's target:
 * pred fp32 [batch, n_chan, h, w]; indices int32 [n_rows, 3] = (b, x, y); target fp32 [n_rows, target_chan] with
 * target_chan = n_chan or 1 (broadcast over channels).  fwd: *loss (device scalar); chunk partials are added in a fixed
 * order.  bwd: d_pred (ZEROED by the caller) gets sign(pred - target) * *grad_out / N at every hit's cell. */
size_t wfsp_segment_l1_workspace_bytes(int64_t n_rows);
int wfsp_segment_l1_fwd(const float* pred, const int32_t* indices, const float* target, int64_t n_rows,
                        const int32_t* n_rows_dev, int n_chan, int target_chan, int batch, int h, int w,
                        float* loss, void* workspace, size_t workspace_bytes, wfsp_stream_t stream);
int wfsp_segment_l1_bwd(const float* pred, const int32_t* indices, const float* target, int64_t n_rows,
                        const int32_t* n_rows_dev, int n_chan, int target_chan, int batch, int h, int w,
                        const float* grad_out, float* d_pred, wfsp_stream_t stream);

/* Column sums of the live rows of x [n_rows, c] (the bias gradient of a convolution on the graph path, where the
 * buffer is capacity-sized and only *n_rows_dev rows are live); workspace: wfsp_bn_workspace_bytes(n_rows, c). */
int wfsp_col_sum(const float* x, int64_t n_rows, const int32_t* n_rows_dev, int c, float* out,
                 void* workspace, size_t workspace_bytes, wfsp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (9) Window edges.  Replaces the reference's only native function, cffi_window_edges
 * (src/custom_functions/cffi.c:5-37, bound in src/custom_functions/__init__.py:5-35 and called from
 * src/utils/GraphUtils.py:7-40): for every hit i an optional self loop, then for every later hit j of the
 * same contiguous run of equal batch ids with |x_i - x_j| < n and |y_i - y_j| < n the edges (i,j), (j,i)
 * (n = max_dist + 1), in exactly the CPU loop's order.  x, y, b: int64 [num_elem]; edges1 / edges2: int64
 * [edge_cap]; *edge_count (device int64) = number of edges.  With edges1 == NULL only the count is
 * produced (first pass to size the output). */
size_t wfsp_window_edges_workspace_bytes(int64_t num_elem);
int wfsp_window_edges(int64_t n, int64_t num_elem, const int64_t* x, const int64_t* y, const int64_t* b,
                      int self_loop, int64_t* edges1, int64_t* edges2, int64_t edge_cap,
                      int64_t* edge_count, void* workspace, size_t workspace_bytes,
                      wfsp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WFSP_H_ */
