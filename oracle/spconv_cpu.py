"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by waveformml_b200/).

A restatement, on the host, of the spconv-1.2.1 CPU path that the reference reaches from
src/models/SPConvBlocks.py / SPConvNet.py: single-threaded hash rulebook
(oracle/wfsp_oracle.c), then per kernel offset  gather -> torch.mm -> scatter-add, and the
matching backward (two GEMMs per offset).  BatchNorm1d / ReLU / Linear stay stock torch-CPU.

PARITY UNPINNED: upstream spconv is a third-party dependency absent from /root/reference
(requirements.txt:15) and the reference has no tests for it; see the header of wfsp_oracle.c.
Anchors that ARE checked (tests/test_oracle.py): SURVEY A.6 known-answer rulebooks and the A.5
dense-convolution identities against torch.nn.functional.conv2d / conv_transpose2d.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
import ctypes
import math
import os

import numpy as np
import torch
from torch import nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libwfsp_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "libwfsp_oracle.so"])
        L = ctypes.CDLL(path)
        i32p, i64p, f32p = (ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64),
                            ctypes.POINTER(ctypes.c_float))
        intp = ctypes.POINTER(ctypes.c_int)
        L.wfo_conv_out_size.restype = ctypes.c_int
        L.wfo_conv_out_size.argtypes = [ctypes.c_int] * 5
        L.wfo_rulebook_conv.restype = ctypes.c_int64
        L.wfo_rulebook_conv.argtypes = [i32p, ctypes.c_int64, ctypes.c_int, intp, intp, intp, intp,
                                        intp, intp, i32p, i32p, i32p]
        L.wfo_rulebook_subm.restype = ctypes.c_int64
        L.wfo_rulebook_subm.argtypes = [i32p, ctypes.c_int64, ctypes.c_int, intp, intp, intp, i32p, i32p]
        L.wfo_rulebook_conv_nd.restype = ctypes.c_int64
        L.wfo_rulebook_conv_nd.argtypes = [ctypes.c_int] + L.wfo_rulebook_conv.argtypes
        L.wfo_rulebook_subm_nd.restype = ctypes.c_int64
        L.wfo_rulebook_subm_nd.argtypes = [ctypes.c_int] + L.wfo_rulebook_subm.argtypes
        L.wfo_gather_rows.restype = None
        L.wfo_gather_rows.argtypes = [f32p, ctypes.c_int64, i32p, ctypes.c_int64, f32p]
        L.wfo_scatter_add_rows.restype = None
        L.wfo_scatter_add_rows.argtypes = [f32p, ctypes.c_int64, i32p, ctypes.c_int64, f32p]
        L.wfo_to_dense.restype = None
        L.wfo_to_dense.argtypes = [f32p, i32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, f32p]
        L.wfo_batch_pack.restype = ctypes.c_int32
        L.wfo_batch_pack.argtypes = [i32p, ctypes.POINTER(ctypes.c_int16), ctypes.c_int64,
                                     ctypes.c_int64, i64p, i64p, ctypes.c_int64, ctypes.c_float,
                                     i32p, f32p]
        L.wfo_window_edges.restype = ctypes.c_int64
        L.wfo_window_edges.argtypes = [ctypes.c_int64, ctypes.c_int64, i64p, i64p, i64p, ctypes.c_int, i64p, i64p]
        _LIB = L
    return _LIB


def window_edges(coo, batch, max_dist=1, self_loops=True):
    """Oracle of src/utils/GraphUtils.py:7-40 (window_edges): coo int64 [N, 2], batch int64 [N] -> int64 [2, E]."""
    coo = np.ascontiguousarray(np.asarray(coo, dtype=np.int64))
    batch = np.ascontiguousarray(np.asarray(batch, dtype=np.int64))
    n = coo.shape[0]
    x, y = np.ascontiguousarray(coo[:, 0]), np.ascontiguousarray(coo[:, 1])
    # bound: every hit pairs with at most (longest run of equal batch ids - 1) later hits
    run = 1
    if n:
        brk = np.flatnonzero(np.diff(batch) != 0)
        run = int(np.max(np.diff(np.concatenate([[-1], brk, [n - 1]]))))
    cap = max(1, n * (1 + 2 * run))
    e1, e2 = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    cnt = lib().wfo_window_edges(max_dist + 1, n, p(x), p(y), p(batch), int(bool(self_loops)), p(e1), p(e2))
    return np.stack([e1[:cnt], e2[:cnt]], 0)


def _p(t, ct):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ct))


def _ints(v, n=2):
    if isinstance(v, (int, np.integer)):
        v = [int(v)] * n
    v = [int(x) for x in v]
    assert len(v) == n
    return (ctypes.c_int * n)(*v)


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    return [lib().wfo_conv_out_size(int(i), int(k), int(s), int(p), int(d))
            for i, k, s, p, d in zip(input_size, kernel_size, stride, padding, dilation)]


def get_indice_pairs(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm=False):
    """Upstream ops.get_indice_pairs, CPU path.  Returns (outids, pairs[2,K,N], pair_num[K]).
    2-d (indices [N,3]) for the 14x11 grid; 3-d (indices [N,4]) for net_type "3DConvolution"."""
    nd = len(spatial_shape)
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == nd + 1 and nd in (2, 3)
    indices = indices.contiguous()
    N = indices.shape[0]
    K = int(np.prod(ksize))
    for s, d in zip(stride, dilation):
        assert s == 1 or d == 1, "don't support this."
    pairs = torch.empty((2, K, N), dtype=torch.int32)
    pair_num = torch.empty((K,), dtype=torch.int32)
    if subm:
        n_out = lib().wfo_rulebook_subm_nd(nd, _p(indices, ctypes.c_int32), N, int(batch_size),
                                           _ints(spatial_shape, nd), _ints(ksize, nd), _ints(dilation, nd),
                                           _p(pairs, ctypes.c_int32), _p(pair_num, ctypes.c_int32))
        assert n_out == N
        return indices, pairs, pair_num
    out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    cap = max(1, min(N * K, int(batch_size) * int(np.prod(out_shape)) if min(out_shape) > 0 else 0))
    outids = torch.empty((cap, nd + 1), dtype=torch.int32)
    n_out = lib().wfo_rulebook_conv_nd(nd, _p(indices, ctypes.c_int32), N, int(batch_size),
                                       _ints(spatial_shape, nd), _ints(out_shape, nd), _ints(ksize, nd),
                                       _ints(stride, nd), _ints(padding, nd), _ints(dilation, nd),
                                       _p(outids, ctypes.c_int32), _p(pairs, ctypes.c_int32),
                                       _p(pair_num, ctypes.c_int32))
    assert n_out >= 0
    return outids[:n_out].clone(), pairs, pair_num


def _gather(feat, idx, n):
    buf = torch.empty((n, feat.shape[1]), dtype=torch.float32)
    lib().wfo_gather_rows(_p(feat, ctypes.c_float), feat.shape[1], _p(idx, ctypes.c_int32), n,
                          _p(buf, ctypes.c_float))
    return buf


def _scatter_add(out, idx, n, buf):
    buf = buf.contiguous()
    lib().wfo_scatter_add_rows(_p(out, ctypes.c_float), out.shape[1], _p(idx, ctypes.c_int32), n,
                               _p(buf, ctypes.c_float))


_OPERAND_ROUNDING = None


def set_operand_rounding(mode):
    """None: plain fp32 (the reference arithmetic).  'bf16': round both GEMM operands to bf16 before
    every product (fp32 accumulate) -- emulates the tensor-core arithmetic of the GPU's bf16 mode so
    a test can separate rounding of the operands from kernel bugs."""
    global _OPERAND_ROUNDING
    assert mode in (None, "bf16")
    _OPERAND_ROUNDING = mode


def _r(t):
    return t.bfloat16().float() if _OPERAND_ROUNDING == "bf16" else t


def indice_conv(features, filters, pairs, pair_num, num_act_out, inverse=False, subm=False):
    """Upstream indiceConv: for each offset gather -> mm -> scatter-add."""
    features = _r(features.contiguous().float())
    K = pairs.shape[1]
    W = _r(filters.reshape(K, filters.shape[-2], filters.shape[-1]).float())
    nums = pair_num.tolist()
    kmax = int(np.argmax(nums)) if K > 0 else 0
    if subm:
        out = torch.mm(features, W[kmax])
    else:
        out = torch.zeros((num_act_out, W.shape[2]), dtype=torch.float32)
    src, dst = (1, 0) if inverse else (0, 1)
    for k in range(K):
        n = nums[k]
        if n <= 0 or (subm and k == kmax):
            continue
        buf = _gather(features, pairs[src, k].contiguous(), n)
        _scatter_add(out, pairs[dst, k].contiguous(), n, torch.mm(buf, W[k]))
    return out


def indice_conv_backward(features, filters, out_bp, pairs, pair_num, inverse=False, subm=False):
    """Upstream indiceConvBackward."""
    features = _r(features.contiguous().float())
    out_bp = _r(out_bp.contiguous().float())
    K = pairs.shape[1]
    W = _r(filters.reshape(K, filters.shape[-2], filters.shape[-1]).float())
    nums = pair_num.tolist()
    kmax = int(np.argmax(nums)) if K > 0 else 0
    dW = torch.zeros_like(W)
    if subm:
        dW[kmax] = torch.mm(features.t(), out_bp)
        dF = torch.mm(out_bp, W[kmax].t())
    else:
        dF = torch.zeros_like(features)
    src, dst = (1, 0) if inverse else (0, 1)
    for k in range(K):
        n = nums[k]
        if n <= 0 or (subm and k == kmax):
            continue
        A = _gather(features, pairs[src, k].contiguous(), n)
        G = _gather(out_bp, pairs[dst, k].contiguous(), n)
        dW[k] = torch.mm(A.t(), G)
        _scatter_add(dF, pairs[src, k].contiguous(), n, torch.mm(G, W[k].t()))
    return dF, dW.reshape(filters.shape)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, filters, pairs, pair_num, num_act_out, inverse, subm):
        ctx.save_for_backward(features, filters, pairs, pair_num)
        ctx.flags = (inverse, subm)
        return indice_conv(features, filters, pairs, pair_num, num_act_out, inverse, subm)

    @staticmethod
    def backward(ctx, grad_output):
        features, filters, pairs, pair_num = ctx.saved_tensors
        inverse, subm = ctx.flags
        dF, dW = indice_conv_backward(features, filters, grad_output, pairs, pair_num, inverse, subm)
        return dF, dW, None, None, None, None, None


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        B, C, nd = self.batch_size, self.features.shape[1], len(self.spatial_shape)
        idx = self.indices.long()
        ret = torch.zeros((B, *self.spatial_shape, C), dtype=self.features.dtype)
        ret[tuple(idx[:, i] for i in range(nd + 1))] = self.features
        if not channels_first:
            return ret
        return ret.permute(0, nd + 1, *range(1, nd + 1)).contiguous()


def to_dense_c(features, indices, batch_size, spatial_shape):
    """wfo_to_dense: the plain-C restatement of .dense() (used to cross-check the torch one)."""
    features = features.contiguous().float()
    indices = indices.contiguous()
    H, W = [int(s) for s in spatial_shape]
    out = torch.zeros((int(batch_size), features.shape[1], H, W), dtype=torch.float32)
    lib().wfo_to_dense(_p(features, ctypes.c_float), _p(indices, ctypes.c_int32), features.shape[0],
                       features.shape[1], int(batch_size), H, W, _p(out, ctypes.c_float))
    return out


class SparseConvolution(nn.Module):
    """Upstream spconv/conv.py SparseConvolution, ndim 2 or 3."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, inverse=False, indice_key=None, ndim=2):
        super().__init__()
        assert groups == 1
        self.ndim = ndim
        t = lambda v: [int(v)] * ndim if isinstance(v, (int, np.integer)) else [int(x) for x in v]
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = t(kernel_size), t(stride), t(padding), t(dilation)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def forward(self, x):
        assert isinstance(x, SparseConvTensor)
        if self.conv1x1:
            if _OPERAND_ROUNDING is None:
                f = torch.mm(x.features, self.weight.view(self.in_channels, self.out_channels))
            else:  # same product through the rounding-aware function (identity pair list)
                n = x.features.shape[0]
                ident = torch.arange(n, dtype=torch.int32).repeat(2, 1, 1)
                f = _ConvFn.apply(x.features, self.weight, ident, torch.tensor([n], dtype=torch.int32), n, False, False)
            if self.bias is not None:
                f = f + self.bias
            out = SparseConvTensor(f, x.indices, x.spatial_shape, x.batch_size)
            out.indice_dict, out.grid = x.indice_dict, x.grid
            return out
        if self.subm:
            out_shape = x.spatial_shape
        else:
            out_shape = get_conv_output_size(x.spatial_shape, self.kernel_size, self.stride, self.padding, self.dilation)
        datas = x.find_indice_pair(self.indice_key)
        if self.inverse:
            assert datas is not None and self.indice_key is not None
            _, outids, pairs, pair_num, out_shape = datas
            assert pairs.shape[1] == int(np.prod(self.kernel_size)), \
                "inverse conv must have same kernel size as its couple conv"
        elif self.indice_key is not None and datas is not None:
            outids, _, pairs, pair_num, _ = datas
        else:
            pad = [k // 2 for k in self.kernel_size] if self.subm else self.padding
            stride = [1] * self.ndim if self.subm else self.stride
            outids, pairs, pair_num = get_indice_pairs(x.indices, x.batch_size, x.spatial_shape,
                                                       self.kernel_size, stride, pad, self.dilation, self.subm)
            x.indice_dict[self.indice_key] = (outids, x.indices, pairs, pair_num, x.spatial_shape)
        f = _ConvFn.apply(x.features, self.weight, pairs, pair_num, outids.shape[0], self.inverse, self.subm)
        if self.bias is not None:
            f = f + self.bias
        out = SparseConvTensor(f, outids, out_shape, x.batch_size)
        out.indice_dict, out.grid = x.indice_dict, x.grid
        return out


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, use_hash=False):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, indice_key=indice_key)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, use_hash=False):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, subm=True, indice_key=indice_key)


class SparseInverseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                         indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, use_hash=False):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, indice_key=indice_key, ndim=3)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, use_hash=False):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, subm=True, indice_key=indice_key, ndim=3)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                         indice_key=indice_key, ndim=3)


class ToDense(nn.Module):
    def forward(self, x):
        return x.dense()


class SparseSequential(nn.Sequential):
    def forward(self, x):
        for m in self:
            if isinstance(m, (SparseConvolution, ToDense)):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x.features = m(x.features)
            else:
                x = m(x)
        return x


def batch_pack(coords, wave, item_rows, item_events, scale):
    """collate_fn + normalisation + batch-first permute (wfo_batch_pack)."""
    coords = coords.contiguous()
    wave = wave.contiguous()
    N, C = wave.shape
    rows = torch.as_tensor(item_rows, dtype=torch.int64).contiguous()
    evs = torch.as_tensor(item_events, dtype=torch.int64).contiguous()
    indices = torch.empty((N, 3), dtype=torch.int32)
    feats = torch.empty((N, C), dtype=torch.float32)
    bs = lib().wfo_batch_pack(_p(coords, ctypes.c_int32), _p(wave, ctypes.c_int16), N, C,
                              _p(rows, ctypes.c_int64), _p(evs, ctypes.c_int64), evs.numel(),
                              ctypes.c_float(scale), _p(indices, ctypes.c_int32), _p(feats, ctypes.c_float))
    return indices, feats, int(bs)
