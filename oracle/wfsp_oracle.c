/*
 * wfsp_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product path (waveformml_b200/, libwfsp.so) never links, imports or
 * calls it.
 *
 * PARITY UNPINNED: the sparse-conv arithmetic of the reference lives in the third-party package
 * spconv~=1.2.1 (/root/reference/requirements.txt:15), which is neither vendored under
 * /root/reference nor installable here, and the reference ships no tests or golden vectors for
 * this path.  This file therefore restates the published algorithm of upstream
 * traveller59/spconv v1.2.1 (CPU path: include/spconv/geometry.h getValidOutPos,
 * src/spconv/indice.cc getIndicePairsConv / getIndicePairsSubM, src/spconv/reordering.cc
 * gather / scatter-add functors, spconv/__init__.py SparseConvTensor.dense) and anchors parity on
 * the reference's own call sites:
 *   - src/models/SPConvNet.py:63-64      (SparseConvTensor construction, batch-first permute)
 *   - src/models/SPConvBlocks.py:498-502 (SparseConv2d positional args nin,nout,fs,st,pd,dil,1,bias)
 *   - src/engineering/PSDDataModule.py:10-20 (collate_fn: running event offset)
 *   - src/datasets/HDF5Dataset.py:15-17,345-346 (normalisation by 1/(2^14-1))
 * plus the hand-derived known-answer vectors of SURVEY.md A.6 and the dense-conv identities of
 * A.5 (tests/test_oracle.py).  The rulebook functions take the number of spatial dimensions at run
 * time (2: the 14x11 grid; 3: net_type "3DConvolution", src/models/SPConvNet.py:42-49); the 3-d case
 * is pinned the same way -- conv3d / conv_transpose3d identities, a hand KAT, rulebook invariants and
 * "a 3-d problem of depth 1 is the 2-d problem" (tests/test_oracle_3d.py).
 *
 * Everything here is single-threaded scalar C on purpose: it is the checker, not the product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WFO_MAX_NDIM 3 /* 2: the 14x11 segment grid; 3: net_type "3DConvolution" (src/models/SPConvNet.py:42-49) */
#define WFO_MAX_KVOL 1024

/* ----------------------------------------------------------------------------------------- */
/* output size: (i + 2p - d(k-1) - 1)/s + 1  (upstream ops.get_conv_output_size; identical to  */
/* the reference's ModelValidation.calc_output_size_1d, src/utils/ModelValidation.py:119-126)  */
int wfo_conv_out_size(int in, int k, int s, int p, int d) {
  int num = in + 2 * p - d * (k - 1) - 1;
  /* python floor division; negative results mean "empty" */
  int q = num / s;
  if ((num % s != 0) && ((num < 0) != (s < 0))) q -= 1;
  return q + 1;
}

/* ----------------------------------------------------------------------------------------- */
/* a tiny open-addressing map int64 -> int32 (stands in for upstream's tsl::robin_map; any     */
/* correct map gives the same rulebook because only insert-if-absent / find are used)          */
typedef struct {
  int64_t* keys;
  int32_t* vals;
  uint64_t mask;
} wfo_map;

static int wfo_map_init(wfo_map* m, int64_t expected) {
  uint64_t cap = 16;
  while ((int64_t)cap < 2 * expected + 2) cap <<= 1;
  m->keys = (int64_t*)malloc(cap * sizeof(int64_t));
  m->vals = (int32_t*)malloc(cap * sizeof(int32_t));
  if (!m->keys || !m->vals) return -1;
  for (uint64_t i = 0; i < cap; ++i) m->keys[i] = -1;
  m->mask = cap - 1;
  return 0;
}
static void wfo_map_free(wfo_map* m) {
  free(m->keys);
  free(m->vals);
}
static inline uint64_t wfo_mix(int64_t k) {
  uint64_t x = (uint64_t)k;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
  return x;
}
/* returns slot of key or of the empty slot where it would go */
static inline uint64_t wfo_map_slot(const wfo_map* m, int64_t key) {
  uint64_t s = wfo_mix(key) & m->mask;
  while (m->keys[s] != -1 && m->keys[s] != key) s = (s + 1) & m->mask;
  return s;
}

/* ----------------------------------------------------------------------------------------- */
/* Candidate output positions of one input point (upstream getValidOutPos).                    */
/* out: rows of (pos[0..D-1], kernel offset).  Returns number of valid rows.                   */
/* Enumeration: last dim fastest, each dim from its upper bound downwards in steps of the      */
/* dilation => ascending kernel offset.  Integer divisions are C divisions (truncate to zero), */
/* exactly as upstream; see SURVEY.md A.2 for why that never changes the emitted pairs.        */
static int wfo_valid_out_pos(int nd, const int32_t* in_pos, const int* ksize, const int* stride,
                             const int* pad, const int* dil, const int* out_shape, int32_t* out) {
  int lowers[WFO_MAX_NDIM], uppers[WFO_MAX_NDIM], counter[WFO_MAX_NDIM], csize[WFO_MAX_NDIM];
  int npoints = 1;
  for (int i = 0; i < nd; ++i) {
    lowers[i] = (in_pos[i] - (ksize[i] - 1) * dil[i] - 1 + stride[i] + pad[i]) / stride[i];
    uppers[i] = (in_pos[i] + pad[i]) / stride[i];
    csize[i] = (uppers[i] - lowers[i]) / dil[i] + 1;
    npoints *= csize[i];
    counter[i] = 0;
  }
  int nvalid = 0;
  for (int i = 0; i < npoints; ++i) {
    int valid = 1, m = 1, offset = 0;
    for (int j = nd - 1; j >= 0; --j) {
      int val = uppers[j] - counter[j] * dil[j];
      out[nvalid * (nd + 1) + j] = val;
      if (val < 0 || val > out_shape[j] - 1) valid = 0;
      offset += m * (in_pos[j] - val * stride[j] + pad[j]) / dil[j];
      m *= ksize[j];
    }
    out[nvalid * (nd + 1) + nd] = offset;
    if (valid) ++nvalid;
    counter[nd - 1] += 1;
    for (int c = nd - 1; c >= 0; --c) {
      if (counter[c] == csize[c] && c > 0) {
        counter[c - 1] += 1;
        counter[c] = 0;
      }
    }
  }
  return nvalid;
}

static inline int64_t wfo_flat(int nd, const int32_t* pos, const int* shape) {
  int64_t f = 0;
  for (int i = 0; i < nd; ++i) f = f * shape[i] + pos[i];
  return f;
}

/* ----------------------------------------------------------------------------------------- */
/* Regular (strided / dilated / padded) conv rulebook, upstream getIndicePairsConv CPU path.   */
/* indices: int32 [N, 1+D] rows (b, x, y[, z]).  pairs: int32 [2, K, N] pre-filled with -1 here. */
/* out_indices must have room for min(N*K, B*vol_out) rows.  Returns N_out, or <0 on error.    */
int64_t wfo_rulebook_conv_nd(int nd, const int32_t* indices, int64_t N, int batch, const int* in_shape,
                             const int* out_shape, const int* ksize, const int* stride,
                             const int* pad, const int* dil, int32_t* out_indices, int32_t* pairs,
                             int32_t* pair_num) {
  (void)batch; (void)in_shape;
  if (nd < 1 || nd > WFO_MAX_NDIM) return -3;
  int K = 1;
  int64_t vol = 1;
  for (int i = 0; i < nd; ++i) { K *= ksize[i]; vol *= out_shape[i]; }
  if (K > WFO_MAX_KVOL) return -2;
  for (int64_t i = 0; i < 2 * (int64_t)K * N; ++i) pairs[i] = -1;
  for (int k = 0; k < K; ++k) pair_num[k] = 0;
  wfo_map map;
  if (wfo_map_init(&map, N * (int64_t)K < (int64_t)batch * vol ? N * (int64_t)K : (int64_t)batch * vol)) return -1;
  int32_t* cand = (int32_t*)malloc((size_t)K * (nd + 1) * sizeof(int32_t));
  int64_t num_act = 0;
  for (int64_t j = 0; j < N; ++j) {
    const int32_t* row = indices + j * (nd + 1);
    int b = row[0];
    int nv = wfo_valid_out_pos(nd, row + 1, ksize, stride, pad, dil, out_shape, cand);
    for (int v = 0; v < nv; ++v) {
      const int32_t* c = cand + v * (nd + 1);
      int off = c[nd];
      int64_t key = wfo_flat(nd, c, out_shape) + vol * (int64_t)b;
      uint64_t s = wfo_map_slot(&map, key);
      if (map.keys[s] == -1) { /* first touch creates the output row */
        map.keys[s] = key;
        map.vals[s] = (int32_t)num_act;
        out_indices[num_act * (nd + 1)] = b;
        for (int i = 0; i < nd; ++i) out_indices[num_act * (nd + 1) + 1 + i] = c[i];
        ++num_act;
      }
      int32_t o = map.vals[s];
      int32_t slot = pair_num[off]++;
      pairs[(0 * (int64_t)K + off) * N + slot] = (int32_t)j;
      pairs[(1 * (int64_t)K + off) * N + slot] = o;
    }
  }
  free(cand);
  wfo_map_free(&map);
  return num_act;
}

/* Submanifold rulebook, upstream getIndicePairsSubM CPU path.  pad := k/2, stride := 1 are    */
/* forced by the caller (upstream spconv_ops.cc).  Output rows == input rows.                  */
int64_t wfo_rulebook_subm_nd(int nd, const int32_t* indices, int64_t N, int batch, const int* shape,
                             const int* ksize, const int* dil, int32_t* pairs, int32_t* pair_num) {
  (void)batch;
  if (nd < 1 || nd > WFO_MAX_NDIM) return -3;
  int K = 1;
  int64_t vol = 1;
  int stride[WFO_MAX_NDIM], pad[WFO_MAX_NDIM];
  for (int i = 0; i < nd; ++i) {
    K *= ksize[i]; vol *= shape[i];
    stride[i] = 1; pad[i] = ksize[i] / 2;
  }
  if (K > WFO_MAX_KVOL) return -2;
  for (int64_t i = 0; i < 2 * (int64_t)K * N; ++i) pairs[i] = -1;
  for (int k = 0; k < K; ++k) pair_num[k] = 0;
  wfo_map map;
  if (wfo_map_init(&map, N)) return -1;
  for (int64_t j = 0; j < N; ++j) { /* later duplicates overwrite */
    const int32_t* row = indices + j * (nd + 1);
    int64_t key = wfo_flat(nd, row + 1, shape) + vol * (int64_t)row[0];
    uint64_t s = wfo_map_slot(&map, key);
    map.keys[s] = key;
    map.vals[s] = (int32_t)j;
  }
  int32_t* cand = (int32_t*)malloc((size_t)K * (nd + 1) * sizeof(int32_t));
  for (int64_t j = 0; j < N; ++j) {
    const int32_t* row = indices + j * (nd + 1);
    int nv = wfo_valid_out_pos(nd, row + 1, ksize, stride, pad, dil, shape, cand);
    for (int v = 0; v < nv; ++v) {
      const int32_t* c = cand + v * (nd + 1);
      int off = c[nd];
      int64_t key = wfo_flat(nd, c, shape) + vol * (int64_t)row[0];
      uint64_t s = wfo_map_slot(&map, key);
      if (map.keys[s] != -1) {
        int32_t slot = pair_num[off]++;
        pairs[(0 * (int64_t)K + off) * N + slot] = (int32_t)j;
        pairs[(1 * (int64_t)K + off) * N + slot] = map.vals[s];
      }
    }
  }
  free(cand);
  wfo_map_free(&map);
  return N;
}

/* the 2-d entry points (the 14x11 grid) */
int64_t wfo_rulebook_conv(const int32_t* indices, int64_t N, int batch, const int* in_shape,
                          const int* out_shape, const int* ksize, const int* stride,
                          const int* pad, const int* dil, int32_t* out_indices, int32_t* pairs,
                          int32_t* pair_num) {
  return wfo_rulebook_conv_nd(2, indices, N, batch, in_shape, out_shape, ksize, stride, pad, dil, out_indices, pairs,
                              pair_num);
}
int64_t wfo_rulebook_subm(const int32_t* indices, int64_t N, int batch, const int* shape,
                          const int* ksize, const int* dil, int32_t* pairs, int32_t* pair_num) {
  return wfo_rulebook_subm_nd(2, indices, N, batch, shape, ksize, dil, pairs, pair_num);
}

/* ----------------------------------------------------------------------------------------- */
/* upstream reordering.cc CPU functors: buf[i,:] = feat[idx[i],:]  /  out[idx[i],:] += buf[i,:] */
void wfo_gather_rows(const float* feat, int64_t C, const int32_t* idx, int64_t n, float* buf) {
  for (int64_t i = 0; i < n; ++i) memcpy(buf + i * C, feat + (int64_t)idx[i] * C, (size_t)C * sizeof(float));
}
void wfo_scatter_add_rows(float* out, int64_t C, const int32_t* idx, int64_t n, const float* buf) {
  for (int64_t i = 0; i < n; ++i) {
    float* o = out + (int64_t)idx[i] * C;
    const float* b = buf + i * C;
    for (int64_t c = 0; c < C; ++c) o[c] += b[c];
  }
}

/* ----------------------------------------------------------------------------------------- */
/* SparseConvTensor.dense(): zeros [B, H, W, C]; ret[b,x,y,:] = features (last write wins);    */
/* then permute to [B, C, H, W] contiguous.  Written directly in NCHW here.                    */
void wfo_to_dense(const float* feat, const int32_t* indices, int64_t N, int64_t C, int batch,
                  int H, int W, float* dense /* [B,C,H,W], caller zero-fills */) {
  (void)batch;
  for (int64_t j = 0; j < N; ++j) {
    const int32_t* r = indices + j * 3;
    for (int64_t c = 0; c < C; ++c)
      dense[(((int64_t)r[0] * C + c) * H + r[1]) * W + r[2]] = feat[j * C + c];
  }
}

/* ----------------------------------------------------------------------------------------- */
/* The batcher: collate_fn (src/engineering/PSDDataModule.py:10-20) + the dtype / scale part   */
/* of HDF5Dataset._concat_range (src/datasets/HDF5Dataset.py:282-346) + the batch-first column  */
/* permute of SPConvNet.forward (src/models/SPConvNet.py:63-64).                               */
/* coords: int32 [N,3] = (x, y, event-local id).  item_rows: [n_items+1] row offsets;          */
/* item_events: [n_items] number of events (labels) in each item.                              */
/* Output: indices int32 [N,3] = (global event, x, y); feats f32 [N,C] = wave * scale.         */
/* Returns batch size = last global event id + 1 (SPConvNet.py:63).                            */
int32_t wfo_batch_pack(const int32_t* coords, const int16_t* wave, int64_t N, int64_t C,
                       const int64_t* item_rows, const int64_t* item_events, int64_t n_items,
                       float scale, int32_t* indices, float* feats) {
  int64_t offset = 0;
  for (int64_t it = 0; it < n_items; ++it) {
    for (int64_t j = item_rows[it]; j < item_rows[it + 1]; ++j) {
      indices[j * 3 + 0] = coords[j * 3 + 2] + (int32_t)(it > 0 ? offset : 0);
      indices[j * 3 + 1] = coords[j * 3 + 0];
      indices[j * 3 + 2] = coords[j * 3 + 1];
    }
    offset += item_events[it];
  }
  for (int64_t i = 0; i < N * C; ++i) feats[i] = (float)wave[i] * scale;
  return N > 0 ? indices[(N - 1) * 3] + 1 : 0;
}

/* ---- window edges -----------------------------------------------------------
 * Restatement of /root/reference/src/custom_functions/cffi.c:5-37 (cffi_window_edges), the reference's only
 * native code (called from src/utils/GraphUtils.py:7-40): for every element i, an optional self loop,
 * then for every later element j of the same contiguous run of equal batch ids with |dx| < n and
 * |dy| < n the two directed edges (i,j), (j,i).  n = max_dist + 1.  Returns the number of edges.
 * (The reference takes abs() of the long long difference through C's int abs; coordinates here are tiny,
 * so the truncation never matters -- noted, not reproduced.)                                          */
int64_t wfo_window_edges(int64_t n, int64_t num_elem, const int64_t* x, const int64_t* y, const int64_t* b,
                         int self_loop, int64_t* edges1, int64_t* edges2) {
  int64_t e = 0;
  for (int64_t i = 0; i < num_elem; ++i) {
    if (self_loop) { edges1[e] = i; edges2[e] = i; ++e; }
    for (int64_t j = i + 1; j < num_elem && b[i] == b[j]; ++j) {
      int64_t dx = x[i] - x[j], dy = y[i] - y[j];
      if (dx < 0) dx = -dx;
      if (dy < 0) dy = -dy;
      if (dx < n && dy < n) {
        edges1[e] = i; edges2[e] = j; ++e;
        edges2[e] = i; edges1[e] = j; ++e;
      }
    }
  }
  return e;
}
