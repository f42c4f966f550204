"""Rulebook ops (mirror of upstream spconv/ops.py: get_conv_output_size, get_indice_pairs) on top
of libwfsp.so.  Reference call sites: every spconv.SparseConv2d / SubMConv2d with kernel volume > 1
in src/models/SPConvBlocks.py (e.g. :498-502)."""
import numpy as np
import torch

from .. import _lib


def _listn(v, nd=2):
    if isinstance(v, (int, np.integer)):
        return [int(v)] * nd
    v = [int(x) for x in v]
    assert len(v) == nd, "expected %d values per geometry argument, got %r" % (nd, v)
    return v


_list2 = _listn


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """(i + 2p - d(k-1) - 1)//s + 1 per dimension; the same arithmetic as the reference's
    ModelValidation.calc_output_size_1d (src/utils/ModelValidation.py:119-126)."""
    out = []
    for i, k, s, p, d in zip(input_size, kernel_size, stride, padding, dilation):
        out.append((int(i) + 2 * int(p) - int(d) * (int(k) - 1) - 1) // int(s) + 1)
    return out


class Rulebook:
    """Geometry of one sparse convolution, cached in SparseConvTensor.indice_dict under the layer's
    indice_key.  Unpacks like upstream's 5-tuple (outids, indices, indice_pairs, indice_pair_num,
    spatial_shape) and additionally carries the output-stationary neighbour tables.

    Graph path: n_in_dev / n_out_dev are int32 device scalars holding the live row counts and every
    tensor is allocated at capacity; eager path: they are None and all shapes are exact."""

    def __init__(self, outids, indices, pairs, pair_num, spatial_shape, out_spatial_shape, nbr_out, nbr_in, dup_flag,
                 n_in_dev=None, n_out_dev=None):
        self.outids, self.indices, self.pairs, self.pair_num = outids, indices, pairs, pair_num
        self.spatial_shape, self.out_spatial_shape = spatial_shape, out_spatial_shape
        self.nbr_out, self.nbr_in, self.dup_flag = nbr_out, nbr_in, dup_flag
        self.n_in_dev, self.n_out_dev = n_in_dev, n_out_dev

    def finish(self):
        """Launches the half of the rulebook that build_rulebook(front_only=True) left for later (no-op otherwise)."""
        pend = getattr(self, "_pending", None)
        while pend:
            pend.pop()()

    def _tuple(self):
        self.finish()
        return (self.outids, self.indices, self.pairs, self.pair_num, self.spatial_shape)

    def __iter__(self):
        return iter(self._tuple())

    def __getitem__(self, i):
        return self._tuple()[i]

    def __len__(self):
        return 5

    @property
    def kvol(self):
        return self.pairs.shape[1]


# Duplicate (batch, x, y[, t]) rows in the input.  Upstream's gather / scatter-add sums all contributors of an
# (output, offset); the output-stationary neighbour tables keep one, so duplicates are an ERROR here, not a silent
# difference: the eager path raises at rulebook construction (the flag rides on the row-count readback), the graph
# path cannot read anything back, so every rulebook built with device-side counts appends its flag to
# `graph_dup_flags` (harness.GraphTrainStep.duplicate_inputs() reads them after a replay).
check_duplicates_default = True
graph_dup_flags = []


def build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm=False,
                   check_duplicates=None, n_rows=None, front_only=False):
    """Builds pairs + neighbour tables on the GPU.

    Eager path (n_rows None): one host readback (n_out) for a regular conv, none for a submanifold
    conv; all results have exact shapes.  Graph path (n_rows = int32 device scalar, indices at
    capacity): no readback at all -- outputs are allocated at their upper bound
    min(N*K, B*H'*W') and the live output count stays on the device.

    front_only (graph path): only the half of the rulebook the forward pass waits for is built now (output rows,
    n_out, nbr_out); the returned Rulebook carries `finish()`, which launches the other half (pairs, pair_num, nbr_in,
    duplicate flag) on the current stream -- any stream ordered after this call -- and must run before backward."""
    lib = _lib.load()
    _lib.require_cuda(indices)
    if indices.dtype != torch.int32:
        raise RuntimeError("indices must be int32 (got %s)" % indices.dtype)
    spatial_shape = [int(s) for s in spatial_shape]
    nd = len(spatial_shape)
    assert nd in (2, 3), "2-d (14x11 grid) and 3-d (14x11xsamples) sparse convolutions are implemented"
    assert indices.dim() == 2 and indices.shape[1] == nd + 1, "indices must be [N, %d] = (batch, x, y%s)" % (
        nd + 1, ", t" if nd == 3 else "")
    indices = indices.contiguous()
    ksize, stride, padding, dilation = (_listn(v, nd) for v in (ksize, stride, padding, dilation))
    for s, d in zip(stride, dilation):
        assert s == 1 or d == 1, "don't support this."
    batch_size = int(batch_size)
    dev = indices.device
    N, K = indices.shape[0], int(np.prod(ksize))
    static = n_rows is not None
    if subm:
        out_shape = list(spatial_shape)
    else:
        out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    pairs = torch.empty((2, K, N), dtype=torch.int32, device=dev)
    pair_num = torch.empty((K,), dtype=torch.int32, device=dev)
    oshape_c = _lib.ints([max(o, 0) for o in out_shape], nd)
    ws_bytes = lib.wfsp_rulebook_workspace_bytes_nd(nd, N, batch_size, oshape_c, _lib.ints(ksize, nd))
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    n_in_dev = n_rows if static else None
    n_out_dev = None
    from .functional import hints
    n_hint = hints.get(n_rows) if static else 0  # expected live rows: picks the builder, never the result
    if check_duplicates is None:
        check_duplicates = check_duplicates_default
    meta = torch.empty((2,), dtype=torch.int32, device=dev)  # [n_out, duplicate flag]: ONE readback on the eager path
    dup = meta[1:2]
    nbr_in = torch.empty((N, K), dtype=torch.int32, device=dev)

    split = bool(front_only and static)
    pending = []  # the BACK half, when it was left for later

    def build(subm_flag, out_indices, out_cap, n_out_t, nbr_out):
        geom = [_lib.ints(v, nd) for v in (spatial_shape, ksize, stride, padding, dilation)]
        tail = (subm_flag, _lib.ptr(out_indices), out_cap, _lib.ptr(pairs), _lib.ptr(pair_num), _lib.ptr(n_out_t),
                _lib.ptr(nbr_out), _lib.ptr(nbr_in), _lib.ptr(dup), _lib.ptr(ws), ws.numel())
        head = (_lib.ptr(indices), N, _lib.ptr(n_in_dev), n_hint, batch_size)
        if split:
            built_all = (_lib.ctypes.c_int * 1)(0)
            rc = lib.wfsp_rulebook_build_phased(nd, *head, *geom, *tail, 1, built_all, _lib.stream())
            if rc == 0 and not built_all[0]:
                def finish():
                    with torch.cuda.device(dev):
                        _lib.check(lib.wfsp_rulebook_build_phased(nd, *head, *geom, *tail, 2, None, _lib.stream()))
                pending.append(finish)
            return rc
        if nd == 2:  # the 14x11 grid keeps its own entry point
            return lib.wfsp_rulebook_build(*head, *geom, *tail, _lib.stream())
        return lib.wfsp_rulebook_build_nd(nd, *head, *geom, *tail, _lib.stream())

    with torch.cuda.device(dev):
        if subm:
            nbr_out = torch.empty((N, K), dtype=torch.int32, device=dev)
            _lib.check(build(1, None, N, None, nbr_out))
            outids, n_out = indices, N
            n_out_dev = n_in_dev
        else:
            cells = batch_size * int(np.prod([max(o, 0) for o in out_shape]))
            cap = max(1, min(N * K, cells))
            outbuf = torch.empty((cap, nd + 1), dtype=torch.int32, device=dev)
            n_out_t = meta[0:1]
            # nbr_out is sized at the bound: its live rows are known only on the device
            nbr_cap = torch.empty((cap, K), dtype=torch.int32, device=dev)
            _lib.check(build(0, outbuf, cap, n_out_t, nbr_cap))
            if static:
                n_out, outids, n_out_dev, nbr_out = cap, outbuf, n_out_t, nbr_cap
            else:
                n_out, dup_host = meta.tolist()  # the one readback per rulebook (output tensor shapes need it)
                outids, nbr_out = outbuf[:n_out], nbr_cap[:n_out]
                if check_duplicates and dup_host != 0:
                    raise RuntimeError(_DUP_MESSAGE)
    if static:
        graph_dup_flags.append(dup)
    elif subm and check_duplicates and N > 0 and int(dup.item()) != 0:
        raise RuntimeError(_DUP_MESSAGE)
    rb = Rulebook(outids, indices, pairs, pair_num, spatial_shape, out_shape, nbr_out, nbr_in, dup,
                  n_in_dev, n_out_dev)
    rb._keep = (ws,)  # the BACK half reuses the workspace
    rb._pending = pending
    return rb


_DUP_MESSAGE = ("duplicate (batch, x, y) coordinates in the input: upstream spconv would sum their contributions, the "
                "output-stationary kernels keep one -- de-duplicate the hits (or sum their features) before the batcher")


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1, out_padding=0,
                     subm=False, transpose=False, grid=None, use_hash=False):
    """Upstream-compatible signature; returns (outids, indice_pairs [2,K,N], indice_pair_num [K])."""
    if transpose:
        raise NotImplementedError("transposed sparse convolution is not used by the reference models")
    nd = len(spatial_shape)
    ks = _listn(ksize, nd)
    pad = [k // 2 for k in ks] if subm else _listn(padding, nd)
    st = [1] * nd if subm else _listn(stride, nd)
    rb = build_rulebook(indices, batch_size, spatial_shape, ks, st, pad, dilation, subm)
    return rb.outids, rb.pairs, rb.pair_num
