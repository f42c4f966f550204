// Rulebook builder for sm_100a: deterministic, bit-exact with the CPU path of upstream
// spconv-1.2.1 get_indice_pairs, which the reference reaches from every
// spconv.SparseConv2d / SubMConv2d with kernel volume > 1 (src/models/SPConvBlocks.py:498-502 ...).
//
// CPU semantics to reproduce on a GPU:
//   * regular conv: walk inputs j in order, their candidate outputs in ascending kernel offset k;
//     the first (j,k) that touches an output cell creates the next output row.  Every candidate
//     gets rank r = j*K + k; phase 1 does atomicMin(rank) into a coordinate table (direct-addressed
//     when batch*out_h*out_w is small, else an open-addressing hash), so the winner of each cell is
//     its first toucher.  Output row id = number of first touchers with a smaller rank = an
//     exclusive scan in (j,k) order (block scan + scan over per-block totals).
//   * pairs of one offset are in ascending input order and each input occurs at most once per
//     offset, so the slot of pair (j,k) = #inputs j' < j valid at k: K independent compactions,
//     done with warp ballots + popc inside a block and the same per-block-total scan across blocks.
//   * submanifold conv: table[cell] = row (largest j wins, "later duplicates overwrite"), then the
//     same per-offset compaction with validity = "neighbour cell is occupied".
//
// HBM traffic is tiny (12 B/row read, 8 B/pair written); these kernels are latency / atomic bound
// and the tables live in L2.
#include "common.cuh"

namespace wfsp {

static int g_force_hash = 0;
void set_force_hash(int v) { g_force_hash = v; }
void set_small_rows(int v);
extern unsigned long long* g_trace;  // conv_umma.cu (wfsp_debug_trace)

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr uint32_t kEmptyKey = 0xffffffffu;
constexpr int kRankInf = 0x7f7f7f7f;  // memset(0x7f) pattern
constexpr int64_t kDirectMaxCells = int64_t(1) << 24;

// Three spatial dimensions (h, w, t).  The 2-d layers of the 14x11 grid are the case in_t = out_t = kt = 1
// with three-column index rows (cols = 3); the 3-d variant (src/models/SPConvNet.py:42-49) has cols = 4.
struct Geom {
  int in_h, in_w, in_t, out_h, out_w, out_t;
  int kh, kw, kt, sh, sw, st, ph, pw, pt, dh, dw, dt;
  int kvol, batch, cols;
  const int32_t* n_dev;  // live row count (graph path); NULL = the host count is exact
  uint32_t mh, mw, mt;   // ceil(2^32 / stride) per dimension (0 for stride 1): division by multiplication
};

// n / s for 0 <= n < 2^32 / s with m = ceil(2^32 / s)  (m == 0: s == 1).  The rulebook kernels of a
// 64-event batch are instruction-fetch bound (a single CTA running cold code once), so the integer
// divisions -- some twenty instructions each when the divisor is a run-time value -- are kept out of them.
__device__ __forceinline__ int fast_div(int n, uint32_t m) { return m ? int(__umulhi(uint32_t(n), m)) : n; }

// every kernel tap in ascending offset order k = (kx * kw + ky) * kt + kz.  Not unrolled on purpose: the
// bodies are long and the compiler would otherwise emit 4 x 4 x 4 copies of each.
#define WFSP_FOR_TAPS(g, kx, ky, kz)                             \
  _Pragma("unroll 1") for (int kx = 0; kx < (g).kh; ++kx)        \
    _Pragma("unroll 1") for (int ky = 0; ky < (g).kw; ++ky)      \
      _Pragma("unroll 1") for (int kz = 0; kz < (g).kt; ++kz)
#define WFSP_TAP(g, kx, ky, kz) (((kx) * (g).kw + (ky)) * (g).kt + (kz))

struct Table {
  int32_t* vals;
  uint32_t* keys;  // hash mode only
  uint32_t mask;   // hash mode only
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <bool HASH>
__device__ __forceinline__ int table_insert(const Table& t, uint32_t key) {
  if (!HASH) return int(key);
  uint32_t s = mix32(key) & t.mask;
  while (true) {
    uint32_t prev = atomicCAS(&t.keys[s], kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) return int(s);
    s = (s + 1) & t.mask;
  }
}

// slot holding `key`, or -1 (hash mode, key absent).  Direct mode always returns the cell itself.
template <bool HASH>
__device__ __forceinline__ int table_find(const Table& t, uint32_t key) {
  if (!HASH) return int(key);
  uint32_t s = mix32(key) & t.mask;
  while (true) {
    uint32_t k = t.keys[s];
    if (k == key) return int(s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & t.mask;
  }
}

// bit kk of the result is set iff kernel tap kk along one dimension maps input coordinate `c`
// onto an existing output coordinate:  c = o*s - p + kk*d  with 0 <= o < out
__device__ __forceinline__ uint32_t tap_mask(int c, int k, int s, uint32_t magic, int p, int d, int out) {
  uint32_t m = 0;
#pragma unroll 1
  for (int kk = 0; kk < k; ++kk) {
    const int num = c + p - kk * d;
    const int q = fast_div(num < 0 ? 0 : num, magic);
    if (num >= 0 && q * s == num && q < out) m |= 1u << kk;
  }
  return m;
}

struct Row {
  int b, x, y, z;
  uint32_t mx, my, mz;
  bool ok;
};

__device__ __forceinline__ Row load_row(const int32_t* __restrict__ indices, int64_t n, int64_t j,
                                        const Geom& g) {
  Row r;
  r.b = r.x = r.y = r.z = 0;
  r.mx = r.my = r.mz = 0;
  // rows up to the capacity n are readable: fetch the coordinates alongside the live count, not after it
  if (j < n) {
    r.b = indices[g.cols * j + 0];
    r.x = indices[g.cols * j + 1];
    r.y = indices[g.cols * j + 2];
    if (g.cols > 3) r.z = indices[g.cols * j + 3];
  }
  r.ok = j < (g.n_dev ? int64_t(*g.n_dev) : n) && j < n;
  if (r.ok) {
    r.ok = r.b >= 0 && r.b < g.batch && r.x >= 0 && r.x < g.in_h && r.y >= 0 && r.y < g.in_w && r.z >= 0 &&
           r.z < g.in_t;
    if (r.ok) {
      r.mx = tap_mask(r.x, g.kh, g.sh, g.mh, g.ph, g.dh, g.out_h);
      r.my = tap_mask(r.y, g.kw, g.sw, g.mw, g.pw, g.dw, g.out_w);
      r.mz = g.kt > 1 ? tap_mask(r.z, g.kt, g.st, g.mt, g.pt, g.dt, g.out_t) : 1u;
    }
  }
  return r;
}

struct Pos {
  int x, y, z;
};

__device__ __forceinline__ void store_pos(int32_t* __restrict__ out_indices, int64_t row, const Geom& g, int b,
                                          const Pos& o) {
  int32_t* p = out_indices + g.cols * row;
  p[0] = b; p[1] = o.x; p[2] = o.y;
  if (g.cols > 3) p[3] = o.z;
}

__device__ __forceinline__ uint32_t in_key(const Geom& g, int b, int x, int y, int z) {
  return uint32_t(((b * g.in_h + x) * g.in_w + y) * g.in_t + z);
}

__device__ __forceinline__ uint32_t out_key(const Row& r, const Geom& g, int kx, int ky, int kz, Pos& o) {
  // only called for taps whose mask bits are set: the numerators are non-negative multiples of the strides
  o.x = fast_div(r.x + g.ph - kx * g.dh, g.mh);
  o.y = fast_div(r.y + g.pw - ky * g.dw, g.mw);
  o.z = fast_div(r.z + g.pt - kz * g.dt, g.mt);
  return uint32_t(((r.b * g.out_h + o.x) * g.out_w + o.y) * g.out_t + o.z);
}

// ---- regular conv, phase 1: atomicMin(rank) per touched output cell ---------------------------
template <bool HASH>
__global__ void __launch_bounds__(kBlock) rb_conv_mark(const int32_t* __restrict__ indices, int64_t n,
                                                       Geom g, Table t) {
  int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  Row r = load_row(indices, n, j, g);
  if (!r.ok) return;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
    if (!(((r.mx >> kx) & 1) && ((r.my >> ky) & 1) && ((r.mz >> kz) & 1))) continue;
    Pos o;
    uint32_t key = out_key(r, g, kx, ky, kz, o);
    int slot = table_insert<HASH>(t, key);
    atomicMin(&t.vals[slot], int(j) * g.kvol + WFSP_TAP(g, kx, ky, kz));
  }
}

// ---- submanifold, phase 1: table[cell] = row (largest row id wins) ----------------------------
template <bool HASH>
__global__ void __launch_bounds__(kBlock) rb_subm_insert(const int32_t* __restrict__ indices, int64_t n,
                                                         Geom g, Table t) {
  int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (j >= (g.n_dev ? int64_t(*g.n_dev) : n)) return;
  Row r = load_row(indices, n, j, g);
  if (!r.ok) return;
  uint32_t key = in_key(g, r.b, r.x, r.y, r.z);
  int slot = table_insert<HASH>(t, key);
  atomicMax(&t.vals[slot], int(j));
}

// validity of candidate (row, kx, ky, kz); for SUBM also returns the output row in `val`,
// for CONV the table slot in `slot`.
template <bool HASH, bool SUBM>
__device__ __forceinline__ bool candidate(const Row& r, const Geom& g, const Table& t, int kx, int ky, int kz,
                                          int& slot, int& val, Pos& o) {
  if (!(r.ok && ((r.mx >> kx) & 1) && ((r.my >> ky) & 1) && ((r.mz >> kz) & 1))) return false;
  uint32_t key = out_key(r, g, kx, ky, kz, o);
  slot = table_find<HASH>(t, key);
  if (SUBM) {
    if (slot < 0) return false;
    val = t.vals[slot];
    return val >= 0;
  }
  return true;
}

// ---- phase 2: per-block totals: K per-offset pair counts (+ first-touch count for CONV) -------
template <bool HASH, bool SUBM>
__global__ void __launch_bounds__(kBlock) rb_count(const int32_t* __restrict__ indices, int64_t n, Geom g,
                                                   Table t, int32_t* __restrict__ blk_cnt) {
  extern __shared__ int s_cnt[];  // kvol + 1
  const int K = g.kvol;
  for (int i = threadIdx.x; i <= K; i += kBlock) s_cnt[i] = 0;
  __syncthreads();
  int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  Row r = load_row(indices, n, j, g);
  const int lane = threadIdx.x & 31;
  int first = 0;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
    int slot = 0, val = 0;
    Pos o;
    bool v = candidate<HASH, SUBM>(r, g, t, kx, ky, kz, slot, val, o);
    const int k = WFSP_TAP(g, kx, ky, kz);
    if (!SUBM && v && t.vals[slot] == int(j) * K + k) ++first;
    unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0 && bal) atomicAdd(&s_cnt[k], __popc(bal));
  }
  if (!SUBM) {
    for (int o = 16; o > 0; o >>= 1) first += __shfl_xor_sync(0xffffffffu, first, o);
    if (lane == 0 && first) atomicAdd(&s_cnt[K], first);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= K; i += kBlock) blk_cnt[int64_t(blockIdx.x) * (K + 1) + i] = s_cnt[i];
}

// ---- phase 3: exclusive scan of every counter across blocks; totals -> pair_num / n_out -------
__global__ void __launch_bounds__(32) rb_scan(const int32_t* __restrict__ blk_cnt, int nblk, int K,
                                              int32_t* __restrict__ blk_base, int32_t* __restrict__ pair_num,
                                              int32_t* __restrict__ n_out) {
  const int c = blockIdx.x;  // counter id, 0..K
  const int lane = threadIdx.x;
  int running = 0;
  for (int base = 0; base < nblk; base += 32) {
    int i = base + lane;
    int v = i < nblk ? blk_cnt[int64_t(i) * (K + 1) + c] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (i < nblk) blk_base[int64_t(i) * (K + 1) + c] = running + incl - v;
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    if (c < K) pair_num[c] = running;
    else if (n_out) *n_out = running;
  }
}

// ---- phase 4 (CONV): first touchers get output row ids in rank order --------------------------
template <bool HASH>
__global__ void __launch_bounds__(kBlock) rb_conv_assign(const int32_t* __restrict__ indices, int64_t n,
                                                         Geom g, Table t, const int32_t* __restrict__ blk_base,
                                                         int32_t* __restrict__ out_indices, int64_t out_cap) {
  __shared__ int s_warp[kWarps];
  const int K = g.kvol;
  int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  Row r = load_row(indices, n, j, g);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int first = 0;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
      int slot = 0, val;
      Pos o;
      if (candidate<HASH, false>(r, g, t, kx, ky, kz, slot, val, o) &&
          t.vals[slot] == int(j) * K + WFSP_TAP(g, kx, ky, kz))
        ++first;
    }
  // block-wide exclusive scan of `first`
  int incl = first;
  for (int o = 1; o < 32; o <<= 1) {
    int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int off = blk_base[int64_t(blockIdx.x) * (K + 1) + K] + incl - first;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  if (first == 0) return;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
      int slot = 0, val;
      Pos o;
      if (candidate<HASH, false>(r, g, t, kx, ky, kz, slot, val, o) &&
          t.vals[slot] == int(j) * K + WFSP_TAP(g, kx, ky, kz)) {
        if (off < out_cap) {
          store_pos(out_indices, off, g, r.b, o);
        }
        t.vals[slot] = -off - 1;  // only the first toucher ever rewrites its cell
        ++off;
      }
    }
}

// ---- phase 5: write the pairs in (offset, ascending input) order ------------------------------
template <bool HASH, bool SUBM>
__global__ void __launch_bounds__(kBlock) rb_pairs(const int32_t* __restrict__ indices, int64_t n, Geom g,
                                                   Table t, const int32_t* __restrict__ blk_base,
                                                   int32_t* __restrict__ pairs) {
  extern __shared__ int s_wcnt[];  // kvol * kWarps
  const int K = g.kvol;
  int64_t j = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  Row r = load_row(indices, n, j, g);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
      int slot = 0, val = 0;
      Pos o;
      bool v = candidate<HASH, SUBM>(r, g, t, kx, ky, kz, slot, val, o);
      unsigned bal = __ballot_sync(0xffffffffu, v);
      if (lane == 0) s_wcnt[(WFSP_TAP(g, kx, ky, kz)) * kWarps + warp] = __popc(bal);
    }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  WFSP_FOR_TAPS(g, kx, ky, kz) {
      const int k = WFSP_TAP(g, kx, ky, kz);
      int slot = 0, val = 0;
      Pos o;
      bool v = candidate<HASH, SUBM>(r, g, t, kx, ky, kz, slot, val, o);
      unsigned bal = __ballot_sync(0xffffffffu, v);
      if (v) {
        int pos = blk_base[int64_t(blockIdx.x) * (K + 1) + k] + __popc(bal & lt);
        for (int w = 0; w < warp; ++w) pos += s_wcnt[k * kWarps + w];
        int o = SUBM ? val : -t.vals[slot] - 1;
        pairs[(int64_t(0) * K + k) * n + pos] = int32_t(j);
        pairs[(int64_t(1) * K + k) * n + pos] = o;
      }
    }
}

// ---- output-stationary tables from the pair lists ---------------------------------------------
__global__ void __launch_bounds__(kBlock) rb_tables(const int32_t* __restrict__ pairs,
                                                    const int32_t* __restrict__ pair_num, int K, int64_t pitch,
                                                    int64_t n_in, int64_t n_out, int32_t* __restrict__ nbr_out,
                                                    int32_t* __restrict__ nbr_in, int32_t* __restrict__ dup_flag) {
  const int k = blockIdx.y;
  int64_t s = int64_t(blockIdx.x) * kBlock + threadIdx.x;
  if (s >= pair_num[k]) return;
  int i = pairs[(int64_t(0) * K + k) * pitch + s];
  int o = pairs[(int64_t(1) * K + k) * pitch + s];
  if (i < 0 || i >= n_in || o < 0 || o >= n_out) { *dup_flag = 2; return; }
  nbr_in[int64_t(i) * K + k] = o;
  int prev = atomicExch(&nbr_out[int64_t(o) * K + k], i);
  if (prev != -1) *dup_flag = 1;
}

// ---- small inputs: the whole rulebook (table, output rows, pairs, neighbour tables) in ONE launch ----
// A single CTA of 1024 threads walks the input rows in rounds of 1024 (thread t = row round*1024 + t).
// Same phases as above with __syncthreads in place of kernel boundaries and running bases carried from
// round to round; the CTA also initialises the coordinate table, the -1 padding of the pair arrays and the
// neighbour tables, so no memset nodes are needed.  The number of rounds follows the LIVE row count, so a
// capacity-sized graph buffer costs nothing.  A 64-event batch of the reference's detector (a few hundred
// hits) is launch-latency bound: this replaces ~11 graph nodes per rulebook by one.
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

constexpr int kSmallBlock = 1024;
constexpr int kSmallWarps = kSmallBlock / 32;
int64_t kSmallMaxRows = 2048;   // live rows (two rounds): beyond this the multi-kernel phases win (wfsp_set_option "rulebook_small_rows")
constexpr int64_t kSmallMaxCap = 65536;    // capacity bound of the single-launch builder (its cost follows the LIVE count)

template <bool SUBM, bool SMEM_TABLE>
__global__ void __launch_bounds__(kSmallBlock) rb_small(const int32_t* __restrict__ indices, int64_t n_cap, Geom g,
                                                        int32_t* __restrict__ table_gmem, int64_t cells,
                                                        int32_t* __restrict__ out_indices, int64_t out_cap,
                                                        int32_t* __restrict__ pairs, int32_t* __restrict__ pair_num,
                                                        int32_t* __restrict__ n_out, int32_t* __restrict__ nbr_out,
                                                        int32_t* __restrict__ nbr_in, int32_t* __restrict__ dup_flag,
                                                        unsigned long long* trace, int phases) {
  // phases: 3 = the whole rulebook; 1 = FRONT only -- output rows + nbr_out, what the forward convolution (and the next
  // rulebook) waits for; 2 = BACK only -- pair lists, nbr_in, pair counts, duplicate check, what backward needs: it
  // rebuilds the coordinate table from the output rows of the front launch and runs beside the forward pass.
#define WFSP_RB_TRACE(slot) do { if (trace != nullptr && threadIdx.x == 0) trace[slot] = (unsigned long long)clock64(); } while (0)
  WFSP_RB_TRACE(0);
  extern __shared__ int s_dyn[];  // [K][kSmallWarps] per-warp pair counts / prefixes, [K] running bases, [cells] table
  __shared__ int s_warp[kSmallWarps];
  __shared__ int s_base;
  const int K = g.kvol;
  int* s_kbase = s_dyn + K * kSmallWarps;
  // the coordinate table lives in shared memory when it fits: every probe is then an on-chip access instead
  // of a dependent L2 round trip (nine per row and phase)
  int32_t* table = SMEM_TABLE ? s_kbase + K : table_gmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Table t{table, nullptr, 0};
  const int n = int(n_cap);
  const int live = g.n_dev ? min(int(*g.n_dev), n) : n;
  const int rounds = (live + kSmallBlock - 1) / kSmallBlock;
  WFSP_RB_TRACE(1);
  // phase 0: initialise.  The -1 padding of the pair arrays is part of the upstream-visible result; with
  // device-side counts it is written for the live rows only (nothing reads the capacity tail).
  // (plain counted loops, not unrolled: this kernel runs once, cold, in one CTA -- its cost is fetching its code)
  const int ncells = int(cells);
  const bool front = (phases & 1) != 0, back = (phases & 2) != 0;
#pragma unroll 1
  for (int i = tid; i < ncells; i += kSmallBlock) table[i] = (SUBM || !front) ? -1 : kRankInf;
  if (back) {
    if (!g.n_dev) {
#pragma unroll 1
      for (int i = tid; i < 2 * K * n; i += kSmallBlock) pairs[i] = -1;
    } else {  // rows [0, live) of every (side, offset) list; the capacity tail stays unspecified
#pragma unroll 1
      for (int c = warp; c < 2 * K; c += kSmallWarps)
#pragma unroll 1
        for (int i = lane; i < live; i += 32) pairs[c * n + i] = -1;
    }
#pragma unroll 1
    for (int i = tid; i < live * K; i += kSmallBlock) nbr_in[i] = -1;
  }
#pragma unroll 1
  for (int k = tid; k < K; k += kSmallBlock) s_kbase[k] = 0;
  if (tid == 0) { if (back) *dup_flag = 0; s_base = 0; }
  __syncthreads();
  WFSP_RB_TRACE(2);
  const Row r0 = load_row(indices, n, tid, g);  // round 0 (all there is up to 1024 rows) reads its row once
  // phase 1: claim cells
  if (!SUBM && !front) {
    // BACK only: the front launch already numbered the output rows; cell of output row r <- -r - 1
    const int n_o = min(int(*n_out), int(out_cap));
#pragma unroll 1
    for (int i = tid; i < n_o; i += kSmallBlock) {
      const int32_t* q = out_indices + g.cols * i;
      table[((q[0] * g.out_h + q[1]) * g.out_w + q[2]) * g.out_t + (g.cols > 3 ? q[3] : 0)] = -i - 1;
    }
  }
#pragma unroll 1
  for (int rd = 0; rd < ((SUBM || front) ? rounds : 0); ++rd) {
    const int j = rd * kSmallBlock + tid;
    Row r = rd == 0 ? r0 : load_row(indices, n, j, g);
    if (!r.ok) continue;
    if (SUBM) {
      atomicMax(&table[in_key(g, r.b, r.x, r.y, r.z)], j);
    } else {
      WFSP_FOR_TAPS(g, kx, ky, kz) {
        if (!(((r.mx >> kx) & 1) && ((r.my >> ky) & 1) && ((r.mz >> kz) & 1))) continue;
        Pos o;
        atomicMin(&table[out_key(r, g, kx, ky, kz, o)], j * K + WFSP_TAP(g, kx, ky, kz));
      }
    }
  }
  __syncthreads();
  WFSP_RB_TRACE(3);
  int64_t rows_out = live;
  if (!SUBM && !front) rows_out = *n_out;
  if (!SUBM && front) {
    // phase 2: first touchers -> output rows in rank order, round after round
#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      const int j = rd * kSmallBlock + tid;
      Row r = rd == 0 ? r0 : load_row(indices, n, j, g);
      // a warp whose 32 rows all lie beyond the live count skips the tap loops (the CTA runs on ONE SM: with a few
      // hundred live rows most of its 32 warps would otherwise spend its issue slots on rows that do not exist)
      const bool warp_live = rd * kSmallBlock + warp * 32 < live;
      int first = 0;
      if (warp_live) {
        WFSP_FOR_TAPS(g, kx, ky, kz) {
          int slot = 0, val;
          Pos o;
          if (candidate<false, false>(r, g, t, kx, ky, kz, slot, val, o) && table[slot] == j * K + WFSP_TAP(g, kx, ky, kz))
            ++first;
        }
      }
      WFSP_RB_TRACE(8);
      const int incl = warp_incl_scan(first, lane);
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      WFSP_RB_TRACE(9);
      // every warp scans the 32 per-warp totals itself (one smem read + five shuffles instead of a serial sum)
      const int wtot = warp_incl_scan(s_warp[lane], lane);
      int off = s_base + incl - first + (warp ? __shfl_sync(0xffffffffu, wtot, warp - 1) : 0);
      const int total = s_base + __shfl_sync(0xffffffffu, wtot, 31);
      // a first toucher only rewrites its own cells, and nobody else's test `== own rank` can succeed on them
      if (first) {
        WFSP_FOR_TAPS(g, kx, ky, kz) {
          int slot = 0, val;
          Pos o;
          if (candidate<false, false>(r, g, t, kx, ky, kz, slot, val, o) && table[slot] == j * K + WFSP_TAP(g, kx, ky, kz)) {
            if (off < out_cap) store_pos(out_indices, off, g, r.b, o);
            table[slot] = -off - 1;
            ++off;
          }
        }
      }
      WFSP_RB_TRACE(10);
      __syncthreads();
      if (tid == 0) s_base = total;
      __syncthreads();
    }
    rows_out = s_base;
    if (tid == 0) *n_out = s_base;
  }
  if (rows_out > out_cap) rows_out = out_cap;
  WFSP_RB_TRACE(4);
  if (front) {
#pragma unroll 1
    for (int i = tid; i < int(rows_out) * K; i += kSmallBlock) nbr_out[i] = -1;
  }
  __syncthreads();
  WFSP_RB_TRACE(5);
  if (!back) {
    // FRONT only: the output-stationary table straight from the coordinate table (no compaction needed for it)
#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      const int j = rd * kSmallBlock + tid;
      if (rd * kSmallBlock + warp * 32 >= live) continue;
      Row r = rd == 0 ? r0 : load_row(indices, n, j, g);
      WFSP_FOR_TAPS(g, kx, ky, kz) {
        int slot = 0, val = 0;
        Pos o;
        if (candidate<false, SUBM>(r, g, t, kx, ky, kz, slot, val, o)) {
          const int o_row = SUBM ? val : -table[slot] - 1;
          if (o_row < out_cap) nbr_out[int64_t(o_row) * K + WFSP_TAP(g, kx, ky, kz)] = j;
        }
      }
    }
    WFSP_RB_TRACE(6);
    return;
  }
  // phase 3: per-offset compaction in ascending input order
  const unsigned lt = (1u << lane) - 1u;
  int dup = 0;  // becomes non-zero if a (row, offset) slot of nbr_out was already taken: duplicate coordinates
#pragma unroll 1
  for (int rd = 0; rd < rounds; ++rd) {
    const int j = rd * kSmallBlock + tid;
    Row r = rd == 0 ? r0 : load_row(indices, n, j, g);
    const bool warp_live = rd * kSmallBlock + warp * 32 < live;  // (see phase 2)
    if (warp_live) {
      WFSP_FOR_TAPS(g, kx, ky, kz) {
        int slot = 0, val = 0;
        Pos o;
        const bool v = candidate<false, SUBM>(r, g, t, kx, ky, kz, slot, val, o);
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (lane == 0) s_dyn[WFSP_TAP(g, kx, ky, kz) * kSmallWarps + warp] = __popc(bal);
      }
    } else {
#pragma unroll 1
      for (int k = lane; k < K; k += 32) s_dyn[k * kSmallWarps + warp] = 0;
    }
    WFSP_RB_TRACE(11);
    __syncthreads();
    WFSP_RB_TRACE(12);
    // warp w turns the 32 per-warp counts of offsets w, w + 32, ... into running positions
#pragma unroll 1
    for (int k = warp; k < K; k += kSmallWarps) {
      const int c = s_dyn[k * kSmallWarps + lane];
      const int incl = warp_incl_scan(c, lane);
      const int base = s_kbase[k];
      __syncwarp();
      s_dyn[k * kSmallWarps + lane] = base + incl - c;
      if (lane == 31) s_kbase[k] = base + incl;
    }
    __syncthreads();
    WFSP_RB_TRACE(13);
    if (warp_live) {
    WFSP_FOR_TAPS(g, kx, ky, kz) {
      const int k = WFSP_TAP(g, kx, ky, kz);
      int slot = 0, val = 0;
      Pos o;
      const bool v = candidate<false, SUBM>(r, g, t, kx, ky, kz, slot, val, o);
      const unsigned bal = __ballot_sync(0xffffffffu, v);
      if (v) {
        const int pos = s_dyn[k * kSmallWarps + warp] + __popc(bal & lt);
        const int o_row = SUBM ? val : -table[slot] - 1;
        pairs[(int64_t(0) * K + k) * n + pos] = j;
        pairs[(int64_t(1) * K + k) * n + pos] = o_row;
        nbr_in[int64_t(j) * K + k] = o_row;
        // the previous occupants are only looked at after the loop: the exchanges of one row overlap
        if (o_row < out_cap) {
          if (front) dup |= atomicExch(&nbr_out[int64_t(o_row) * K + k], j) + 1;
          else dup |= int(nbr_out[int64_t(o_row) * K + k] != j);  // the front launch's table: another row took the slot
        }
      }
    }
    }
    WFSP_RB_TRACE(14);
    __syncthreads();
  }
  WFSP_RB_TRACE(6);
  if (dup) *dup_flag = 1;
#pragma unroll 1
  for (int k = tid; k < K; k += kSmallBlock) pair_num[k] = s_kbase[k];
  WFSP_RB_TRACE(7);
#undef WFSP_RB_TRACE
}

struct Plan {
  bool hash;
  int64_t table_slots;
  int nblk;
  size_t off_vals, off_keys, off_cnt, off_base, total;
};

Plan make_plan(int64_t n_in, int batch, int64_t out_vol, int kvol) {
  Plan p;
  int64_t cells = int64_t(batch) * out_vol;
  if (cells < 1) cells = 1;
  p.hash = g_force_hash || cells > kDirectMaxCells;
  if (p.hash) {
    int64_t want = n_in * int64_t(kvol);
    if (want > cells) want = cells;
    int64_t cap = 64;
    while (cap < 2 * want + 2) cap <<= 1;
    p.table_slots = cap;
  } else {
    p.table_slots = cells;
  }
  p.nblk = int(ceil_div<int64_t>(n_in > 0 ? n_in : 1, kBlock));
  size_t o = 0;
  p.off_vals = o; o += align_up(size_t(p.table_slots) * 4, 256);
  p.off_keys = o; o += p.hash ? align_up(size_t(p.table_slots) * 4, 256) : 0;
  p.off_cnt = o;  o += align_up(size_t(p.nblk) * (kvol + 1) * 4, 256);
  p.off_base = o; o += align_up(size_t(p.nblk) * (kvol + 1) * 4, 256);
  p.total = o;
  return p;
}

int check_geom(int ndim, const int* ksize, const int* stride, const int* pad, const int* dil) {
  WFSP_REQUIRE(ndim == 2 || ndim == 3, "ndim %d: only 2-d and 3-d rulebooks", ndim);
  for (int i = 0; i < ndim; ++i) {
    WFSP_REQUIRE(ksize[i] >= 1 && ksize[i] <= 32, "kernel size %d outside [1,32]", ksize[i]);
    WFSP_REQUIRE(stride[i] >= 1 && dil[i] >= 1 && pad[i] >= 0, "bad stride/dilation/padding");
    WFSP_REQUIRE(stride[i] == 1 || dil[i] == 1, "don't support this: stride>1 with dilation>1");
  }
  return WFSP_OK;
}

int out_shape_nd(int ndim, const int* in_shape, const int* ksize, const int* stride, const int* pad, const int* dil,
                 int* out_shape) {
  for (int i = 0; i < ndim; ++i) {
    WFSP_REQUIRE(stride[i] >= 1, "stride must be >= 1");
    int num = in_shape[i] + 2 * pad[i] - dil[i] * (ksize[i] - 1) - 1;
    int q = num >= 0 ? num / stride[i] : -((-num + stride[i] - 1) / stride[i]);  // floor
    out_shape[i] = q + 1;
  }
  return WFSP_OK;
}

// the geometry of one layer, padded to three dimensions (a 2-d layer has a third dimension of extent 1)
struct Geom3 {
  int in[3], out[3], k[3], s[3], p[3], d[3];
  int kvol;
  int64_t out_vol;
};

Geom3 pad3(int ndim, const int* in_shape, const int* out_shape, const int* ksize, const int* stride, const int* pad,
           const int* dil) {
  Geom3 q;
  q.kvol = 1;
  q.out_vol = 1;
  for (int i = 0; i < 3; ++i) {
    const bool on = i < ndim;
    q.in[i] = on ? in_shape[i] : 1;
    q.out[i] = on ? (out_shape[i] > 0 ? out_shape[i] : 0) : 1;
    q.k[i] = on ? ksize[i] : 1;
    q.s[i] = on ? stride[i] : 1;
    q.p[i] = on ? pad[i] : 0;
    q.d[i] = on ? dil[i] : 1;
    q.kvol *= q.k[i];
    q.out_vol *= q.out[i];
  }
  return q;
}

Geom make_geom(const Geom3& q, int ndim, int batch, const int32_t* n_dev) {
  uint32_t magic[3];
  for (int i = 0; i < 3; ++i) magic[i] = q.s[i] > 1 ? uint32_t(((uint64_t(1) << 32) + q.s[i] - 1) / q.s[i]) : 0u;
  return Geom{q.in[0], q.in[1], q.in[2], q.out[0], q.out[1], q.out[2], q.k[0], q.k[1], q.k[2], q.s[0], q.s[1], q.s[2],
              q.p[0], q.p[1], q.p[2], q.d[0], q.d[1], q.d[2], q.kvol, batch, ndim + 1, n_dev, magic[0], magic[1],
              magic[2]};
}

template <bool HASH>
int run_conv(const int32_t* indices, int64_t n, const Geom& g, const Table& t, const Plan& p, char* ws,
             int32_t* out_indices, int64_t out_cap, int32_t* pairs, int32_t* pair_num, int32_t* n_out,
             cudaStream_t st) {
  const int K = g.kvol;
  int32_t* blk_cnt = reinterpret_cast<int32_t*>(ws + p.off_cnt);
  int32_t* blk_base = reinterpret_cast<int32_t*>(ws + p.off_base);
  rb_conv_mark<HASH><<<p.nblk, kBlock, 0, st>>>(indices, n, g, t);
  rb_count<HASH, false><<<p.nblk, kBlock, (K + 1) * sizeof(int), st>>>(indices, n, g, t, blk_cnt);
  rb_scan<<<K + 1, 32, 0, st>>>(blk_cnt, p.nblk, K, blk_base, pair_num, n_out);
  rb_conv_assign<HASH><<<p.nblk, kBlock, 0, st>>>(indices, n, g, t, blk_base, out_indices, out_cap);
  rb_pairs<HASH, false><<<p.nblk, kBlock, K * kWarps * sizeof(int), st>>>(indices, n, g, t, blk_base, pairs);
  count_launches(5);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

template <bool HASH>
int run_subm(const int32_t* indices, int64_t n, const Geom& g, const Table& t, const Plan& p, char* ws,
             int32_t* pairs, int32_t* pair_num, cudaStream_t st) {
  const int K = g.kvol;
  int32_t* blk_cnt = reinterpret_cast<int32_t*>(ws + p.off_cnt);
  int32_t* blk_base = reinterpret_cast<int32_t*>(ws + p.off_base);
  rb_subm_insert<HASH><<<p.nblk, kBlock, 0, st>>>(indices, n, g, t);
  rb_count<HASH, true><<<p.nblk, kBlock, (K + 1) * sizeof(int), st>>>(indices, n, g, t, blk_cnt);
  rb_scan<<<K, 32, 0, st>>>(blk_cnt, p.nblk, K, blk_base, pair_num, nullptr);
  rb_pairs<HASH, true><<<p.nblk, kBlock, K * kWarps * sizeof(int), st>>>(indices, n, g, t, blk_base, pairs);
  count_launches(4);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

int rulebook_conv_impl(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                       const int* in_shape, const int* ksize, const int* stride, const int* pad, const int* dil,
                       int32_t* out_indices, int64_t out_cap, int32_t* pairs, int32_t* pair_num, int32_t* n_out,
                       void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  if (int rc = check_geom(ndim, ksize, stride, pad, dil)) return rc;
  int out_shape[3];
  out_shape_nd(ndim, in_shape, ksize, stride, pad, dil, out_shape);
  WFSP_REQUIRE(n_in >= 0 && batch >= 0, "negative sizes");
  const Geom3 q = pad3(ndim, in_shape, out_shape, ksize, stride, pad, dil);
  const Geom g = make_geom(q, ndim, batch, n_in_dev);
  WFSP_REQUIRE(g.kvol <= WFSP_MAX_KVOL, "kernel volume %d > %d", g.kvol, WFSP_MAX_KVOL);
  WFSP_REQUIRE(n_in * int64_t(g.kvol) < int64_t(kRankInf), "n_in * kvol too large for 31-bit ranks");
  WFSP_REQUIRE(int64_t(batch) * q.out_vol < int64_t(0xfffffff0u), "batch * output volume too large");
  cudaStream_t st = as_stream(stream);
  Plan p = make_plan(n_in, batch, q.out_vol, g.kvol);
  if (workspace_bytes < p.total) return set_error(WFSP_EWORKSPACE, "rulebook workspace %zu < %zu", workspace_bytes, p.total);
  char* ws = static_cast<char*>(workspace);
  Table t{reinterpret_cast<int32_t*>(ws + p.off_vals), reinterpret_cast<uint32_t*>(ws + p.off_keys),
          uint32_t(p.table_slots - 1)};
  WFSP_CHECK_CUDA(cudaMemsetAsync(t.vals, 0x7f, size_t(p.table_slots) * 4, st));
  if (p.hash) WFSP_CHECK_CUDA(cudaMemsetAsync(t.keys, 0xff, size_t(p.table_slots) * 4, st));
  if (n_in > 0) WFSP_CHECK_CUDA(cudaMemsetAsync(pairs, 0xff, size_t(2) * g.kvol * n_in * 4, st));
  if (n_in == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(pair_num, 0, size_t(g.kvol) * 4, st));
    WFSP_CHECK_CUDA(cudaMemsetAsync(n_out, 0, 4, st));
    return WFSP_OK;
  }
  return p.hash ? run_conv<true>(indices, n_in, g, t, p, ws, out_indices, out_cap, pairs, pair_num, n_out, st)
                : run_conv<false>(indices, n_in, g, t, p, ws, out_indices, out_cap, pairs, pair_num, n_out, st);
}

int rulebook_subm_impl(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                       const int* shape, const int* ksize, const int* dil, int32_t* pairs, int32_t* pair_num,
                       void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  const int one[3] = {1, 1, 1};
  const int pad[3] = {ksize[0] / 2, ksize[1] / 2, ndim > 2 ? ksize[2] / 2 : 0};
  if (int rc = check_geom(ndim, ksize, one, pad, dil)) return rc;
  WFSP_REQUIRE(n_in >= 0 && batch >= 0, "negative sizes");
  const Geom3 q = pad3(ndim, shape, shape, ksize, one, pad, dil);
  const Geom g = make_geom(q, ndim, batch, n_in_dev);
  WFSP_REQUIRE(g.kvol <= WFSP_MAX_KVOL, "kernel volume %d > %d", g.kvol, WFSP_MAX_KVOL);
  WFSP_REQUIRE(int64_t(batch) * q.out_vol < int64_t(0xfffffff0u), "batch * volume too large");
  cudaStream_t st = as_stream(stream);
  Plan p = make_plan(n_in, batch, q.out_vol, g.kvol);
  if (workspace_bytes < p.total) return set_error(WFSP_EWORKSPACE, "rulebook workspace %zu < %zu", workspace_bytes, p.total);
  char* ws = static_cast<char*>(workspace);
  Table t{reinterpret_cast<int32_t*>(ws + p.off_vals), reinterpret_cast<uint32_t*>(ws + p.off_keys),
          uint32_t(p.table_slots - 1)};
  WFSP_CHECK_CUDA(cudaMemsetAsync(t.vals, 0xff, size_t(p.table_slots) * 4, st));
  if (p.hash) WFSP_CHECK_CUDA(cudaMemsetAsync(t.keys, 0xff, size_t(p.table_slots) * 4, st));
  if (n_in == 0) {
    WFSP_CHECK_CUDA(cudaMemsetAsync(pair_num, 0, size_t(g.kvol) * 4, st));
    return WFSP_OK;
  }
  WFSP_CHECK_CUDA(cudaMemsetAsync(pairs, 0xff, size_t(2) * g.kvol * n_in * 4, st));
  return p.hash ? run_subm<true>(indices, n_in, g, t, p, ws, pairs, pair_num, st)
                : run_subm<false>(indices, n_in, g, t, p, ws, pairs, pair_num, st);
}

}  // namespace
}  // namespace wfsp

using namespace wfsp;

extern "C" int wfsp_conv_out_shape(const int* in_shape, const int* ksize, const int* stride, const int* pad,
                                   const int* dil, int* out_shape) {
  return out_shape_nd(2, in_shape, ksize, stride, pad, dil, out_shape);
}

extern "C" int wfsp_conv_out_shape_nd(int ndim, const int* in_shape, const int* ksize, const int* stride,
                                      const int* pad, const int* dil, int* out_shape) {
  WFSP_REQUIRE(ndim == 2 || ndim == 3, "ndim %d: only 2-d and 3-d", ndim);
  return out_shape_nd(ndim, in_shape, ksize, stride, pad, dil, out_shape);
}

extern "C" size_t wfsp_rulebook_workspace_bytes(int64_t n_in, int batch, const int* out_shape, const int* ksize) {
  return make_plan(n_in, batch, int64_t(out_shape[0]) * out_shape[1], ksize[0] * ksize[1]).total;
}

extern "C" size_t wfsp_rulebook_workspace_bytes_nd(int ndim, int64_t n_in, int batch, const int* out_shape,
                                                   const int* ksize) {
  int64_t vol = 1;
  int kvol = 1;
  for (int i = 0; i < ndim && i < 3; ++i) { vol *= out_shape[i] > 0 ? out_shape[i] : 0; kvol *= ksize[i]; }
  return make_plan(n_in, batch, vol, kvol).total;
}

extern "C" int wfsp_rulebook_conv(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                                  const int* in_shape,
                                  const int* ksize, const int* stride, const int* pad, const int* dil,
                                  int32_t* out_indices, int64_t out_cap, int32_t* pairs, int32_t* pair_num,
                                  int32_t* n_out, void* workspace, size_t workspace_bytes,
                                  wfsp_stream_t stream) {
  return rulebook_conv_impl(2, indices, n_in, n_in_dev, batch, in_shape, ksize, stride, pad, dil, out_indices, out_cap,
                            pairs, pair_num, n_out, workspace, workspace_bytes, stream);
}

extern "C" int wfsp_rulebook_subm(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int batch,
                                  const int* shape,
                                  const int* ksize, const int* dil, int32_t* pairs, int32_t* pair_num,
                                  void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return rulebook_subm_impl(2, indices, n_in, n_in_dev, batch, shape, ksize, dil, pairs, pair_num, workspace,
                            workspace_bytes, stream);
}

extern "C" int wfsp_rulebook_tables(const int32_t* pairs, const int32_t* pair_num, int kvol, int64_t pair_pitch,
                                    int64_t n_in, int64_t n_out, int32_t* nbr_out, int32_t* nbr_in,
                                    int32_t* dup_flag, wfsp_stream_t stream) {
  WFSP_REQUIRE(kvol >= 1 && kvol <= WFSP_MAX_KVOL, "kvol %d out of range", kvol);
  cudaStream_t st = as_stream(stream);
  if (n_out > 0) WFSP_CHECK_CUDA(cudaMemsetAsync(nbr_out, 0xff, size_t(n_out) * kvol * 4, st));
  if (n_in > 0) WFSP_CHECK_CUDA(cudaMemsetAsync(nbr_in, 0xff, size_t(n_in) * kvol * 4, st));
  if (pair_pitch == 0 || n_in == 0 || n_out == 0) return WFSP_OK;
  dim3 grid(unsigned(ceil_div<int64_t>(pair_pitch, kBlock)), unsigned(kvol));
  rb_tables<<<grid, kBlock, 0, st>>>(pairs, pair_num, kvol, pair_pitch, n_in, n_out, nbr_out, nbr_in, dup_flag);
  count_launches(1);
  WFSP_CHECK_LAUNCH();
  return WFSP_OK;
}

// Rulebook + neighbour tables in one call (what a layer needs before its first convolution).  Small
// inputs (up to kSmallMaxRows expected live rows, direct table, kernel volume <= 256) take the single-launch
// path; everything else runs the phase kernels above followed by rb_tables.  ndim 2: index rows (b, x, y);
// ndim 3: (b, x, y, t).
namespace wfsp { void set_small_rows(int v) { kSmallMaxRows = v; } }

extern "C" int wfsp_rulebook_build_phased(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev,
                                          int64_t n_in_hint, int batch, const int* in_shape, const int* ksize,
                                          const int* stride, const int* pad, const int* dil, int subm,
                                          int32_t* out_indices, int64_t out_cap, int32_t* pairs, int32_t* pair_num,
                                          int32_t* n_out, int32_t* nbr_out, int32_t* nbr_in, int32_t* dup_flag,
                                          void* workspace, size_t workspace_bytes, int phases, int* built_all,
                                          wfsp_stream_t stream) {
  WFSP_REQUIRE(ndim == 2 || ndim == 3, "ndim %d: only 2-d and 3-d rulebooks", ndim);
  WFSP_REQUIRE(phases >= 1 && phases <= 3, "phases must be 1 (front), 2 (back) or 3 (all)");
  if (built_all) *built_all = 0;
  WFSP_REQUIRE(pair_num && dup_flag, "null output");
  WFSP_REQUIRE(n_in == 0 || (pairs && nbr_out && nbr_in), "null output");
  WFSP_REQUIRE(subm || n_out, "regular convolution needs n_out");
  const int one[3] = {1, 1, 1};
  const int spad[3] = {ksize[0] / 2, ksize[1] / 2, ndim > 2 ? ksize[2] / 2 : 0};
  const int* st_ = subm ? one : stride;
  const int* pd_ = subm ? spad : pad;
  if (int rc = check_geom(ndim, ksize, st_, pd_, dil)) return rc;
  int out_shape[3] = {in_shape[0], in_shape[1], ndim > 2 ? in_shape[2] : 1};
  if (!subm) out_shape_nd(ndim, in_shape, ksize, stride, pad, dil, out_shape);
  const Geom3 q = pad3(ndim, in_shape, out_shape, ksize, st_, pd_, dil);
  const int kvol = q.kvol;
  WFSP_REQUIRE(kvol <= WFSP_MAX_KVOL, "kernel volume %d > %d", kvol, WFSP_MAX_KVOL);
  const int64_t cells = int64_t(batch) * q.out_vol;
  cudaStream_t st = as_stream(stream);
  // expected live rows: the capacity, or the caller's hint where only the device knows the count (graph path);
  // a wrong hint only costs time (the single-launch builder walks the live rows in rounds of 1024)
  const int64_t live = (n_in_dev != nullptr && n_in_hint > 0 && n_in_hint < n_in) ? n_in_hint : n_in;
  if (n_in > 0 && live <= kSmallMaxRows && n_in <= kSmallMaxCap && kvol <= 256 && !g_force_hash && cells > 0 && cells <= (int64_t(1) << 20) &&
      n_in * int64_t(kvol) < int64_t(kRankInf)) {
    WFSP_REQUIRE(workspace_bytes >= size_t(cells > 0 ? cells : 1) * 4, "rulebook workspace too small");
    const Geom g = make_geom(q, ndim, batch, n_in_dev);
    size_t smem = size_t(kvol) * (kSmallWarps + 1) * sizeof(int);
    const bool in_smem = smem + size_t(cells) * 4 <= size_t(200) * 1024;
    if (in_smem) smem += size_t(cells) * 4;
    int32_t* tab = static_cast<int32_t*>(workspace);
#define WFSP_RB_SMALL(SUBM, SM, OUT, CAP, NOUT)                                                                  \
    do {                                                                                                        \
      WFSP_CHECK_CUDA(cudaFuncSetAttribute(rb_small<SUBM, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                           208 * 1024));                                                        \
      rb_small<SUBM, SM><<<1, kSmallBlock, smem, st>>>(indices, n_in, g, tab, cells, OUT, CAP, pairs, pair_num,  \
                                                       NOUT, nbr_out, nbr_in, dup_flag, g_trace, phases);        \
    } while (0)
    if (subm) {
      if (in_smem) WFSP_RB_SMALL(true, true, nullptr, n_in, nullptr);
      else WFSP_RB_SMALL(true, false, nullptr, n_in, nullptr);
    } else {
      if (in_smem) WFSP_RB_SMALL(false, true, out_indices, out_cap, n_out);
      else WFSP_RB_SMALL(false, false, out_indices, out_cap, n_out);
    }
#undef WFSP_RB_SMALL
    count_launches(1);
    WFSP_CHECK_LAUNCH();
    return WFSP_OK;
  }
  // the multi-kernel builders are not split: a FRONT request builds everything (and says so), a BACK request after
  // that has nothing left to do
  if (phases == 2) return WFSP_OK;
  if (built_all) *built_all = 1;
  int rc;
  if (subm)
    rc = rulebook_subm_impl(ndim, indices, n_in, n_in_dev, batch, in_shape, ksize, dil, pairs, pair_num, workspace,
                            workspace_bytes, stream);
  else
    rc = rulebook_conv_impl(ndim, indices, n_in, n_in_dev, batch, in_shape, ksize, stride, pad, dil, out_indices,
                            out_cap, pairs, pair_num, n_out, workspace, workspace_bytes, stream);
  if (rc) return rc;
  WFSP_CHECK_CUDA(cudaMemsetAsync(dup_flag, 0, 4, st));
  return wfsp_rulebook_tables(pairs, pair_num, kvol, n_in, n_in, subm ? n_in : out_cap, nbr_out, nbr_in, dup_flag, stream);
}

extern "C" int wfsp_rulebook_build_nd(int ndim, const int32_t* indices, int64_t n_in, const int32_t* n_in_dev,
                                      int64_t n_in_hint, int batch, const int* in_shape, const int* ksize,
                                      const int* stride, const int* pad, const int* dil, int subm,
                                      int32_t* out_indices, int64_t out_cap, int32_t* pairs, int32_t* pair_num,
                                      int32_t* n_out, int32_t* nbr_out, int32_t* nbr_in, int32_t* dup_flag,
                                      void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_rulebook_build_phased(ndim, indices, n_in, n_in_dev, n_in_hint, batch, in_shape, ksize, stride, pad, dil, subm,
                                    out_indices, out_cap, pairs, pair_num, n_out, nbr_out, nbr_in, dup_flag, workspace,
                                    workspace_bytes, 3, nullptr, stream);
}

extern "C" int wfsp_rulebook_build(const int32_t* indices, int64_t n_in, const int32_t* n_in_dev, int64_t n_in_hint, int batch,
                                   const int* in_shape, const int* ksize, const int* stride, const int* pad,
                                   const int* dil, int subm, int32_t* out_indices, int64_t out_cap, int32_t* pairs,
                                   int32_t* pair_num, int32_t* n_out, int32_t* nbr_out, int32_t* nbr_in,
                                   int32_t* dup_flag, void* workspace, size_t workspace_bytes, wfsp_stream_t stream) {
  return wfsp_rulebook_build_nd(2, indices, n_in, n_in_dev, n_in_hint, batch, in_shape, ksize, stride, pad, dil, subm,
                                out_indices, out_cap, pairs, pair_num, n_out, nbr_out, nbr_in, dup_flag, workspace,
                                workspace_bytes, stream);
}
