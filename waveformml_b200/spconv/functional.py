"""Autograd functions over libwfsp.so (mirror of upstream spconv/functional.py:
indice_conv / indice_subm_conv / indice_inverse_conv and the dense scatter).

Every function takes optional int32 device scalars with the live row counts (graph path, see
include/wfsp.h "DEVICE-SIDE ROW COUNTS"); with None the tensor shapes are exact (eager path)."""
import torch
from torch.autograd import Function

from .. import _lib

_MATH = {"fp32": _lib.MATH_FP32, "bf16": _lib.MATH_BF16, "bf16x3": _lib.MATH_BF16X3}


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class LaunchHints:
    """Graph path only.  The host never learns the live row counts there, yet launch shapes (how the
    columns are tiled, how the wgrad reduction is split) should follow them rather than the buffer
    capacities.  During one eager warm-up pass (`record`) every kernel call reads its live count back
    and appends it; the captured pass (`replay`) consumes the same numbers in the same call order.
    They only shape launches; correctness never depends on them."""

    def __init__(self):
        self.mode, self.values, self.pos = None, [], 0

    def start(self, mode):
        self.mode, self.pos = mode, 0
        if mode == "record":
            self.values = []

    def stop(self):
        self.mode = None

    def get(self, count):
        """count: device tensor (its max is taken), or None."""
        if count is None or self.mode is None:
            return 0
        if self.mode == "record":
            v = int(count.max().item())
            self.values.append(v)
            return v
        v = self.values[self.pos] if self.pos < len(self.values) else 0
        self.pos += 1
        return v


hints = LaunchHints()


def live_col_sum(x, n_rows_dev):
    """Column sums over the LIVE rows of a capacity-sized [cap, c] fp32 buffer (wfsp_col_sum): the bias gradient of a
    convolution on the graph path -- a masked torch reduction would stream the whole capacity (157 696 x 150 floats for
    the z-regression model at 1024 events: 80 us for 3 000 live rows)."""
    lib = _lib.load()
    x = x.contiguous()
    n, c = x.shape
    out = torch.empty((c,), dtype=torch.float32, device=x.device)
    ws = torch.empty((lib.wfsp_bn_workspace_bytes(max(n, 1), c),), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.wfsp_col_sum(_lib.ptr(x), n, _lib.ptr(n_rows_dev), c, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                    _lib.stream()))
    return out


def conv_apply(src, weight3, transpose_w, bias, nbr, n_dst, c_dst, math, n_src_dev=None, n_dst_dev=None):
    """dst[r] = bias + sum_k src[nbr[r,k]] @ (weight3[k] or weight3[k]^T)   (wfsp_conv_apply)"""
    lib = _lib.load()
    _lib.require_cuda(src, weight3)
    kvol = weight3.shape[0]
    c_red = src.shape[1]
    dst = torch.empty((n_dst, c_dst), dtype=torch.float32, device=src.device)
    if n_dst == 0:
        return dst
    m = _MATH[math]
    hint = hints.get(n_dst_dev)
    ws_bytes = lib.wfsp_conv_apply_workspace_bytes(kvol, src.shape[0], c_red, c_dst, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=src.device) if ws_bytes else None
    with torch.cuda.device(src.device):
        _lib.check(lib.wfsp_conv_apply(_lib.ptr(src), src.shape[0], _lib.ptr(n_src_dev), c_red, _lib.ptr(weight3),
                                       int(transpose_w), _lib.ptr(bias), _lib.ptr(nbr), kvol, _lib.ptr(dst), n_dst,
                                       _lib.ptr(n_dst_dev), hint, c_dst, m, _lib.ptr(ws), ws_bytes, _lib.stream()))
    return dst


def conv_wgrad(a, b, pair_a, pair_b, pair_num, kvol, math, n_a_dev=None, n_b_dev=None):
    """d_weight[k] = sum_p a[pair_a[k,p]]^T (x) b[pair_b[k,p]]   (wfsp_conv_wgrad)"""
    lib = _lib.load()
    c_a, c_b = a.shape[1], b.shape[1]
    dw = torch.empty((kvol, c_a, c_b), dtype=torch.float32, device=a.device)
    pitch = pair_a.shape[-1] if pair_a is not None else a.shape[0]
    m = _MATH[math]
    ws_bytes = lib.wfsp_conv_wgrad_workspace_bytes(kvol, a.shape[0], c_a, b.shape[0], c_b, pitch, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=a.device) if ws_bytes else None
    hint = 0
    if n_a_dev is not None:  # graph path: expected pairs of the fullest offset
        hint = hints.get(pair_num if pair_num is not None else n_a_dev)
    with torch.cuda.device(a.device):
        _lib.check(lib.wfsp_conv_wgrad(_lib.ptr(a), a.shape[0], _lib.ptr(n_a_dev), c_a, _lib.ptr(b), b.shape[0],
                                       _lib.ptr(n_b_dev), c_b, _lib.ptr(pair_a), _lib.ptr(pair_b), _lib.ptr(pair_num),
                                       kvol, pitch, hint, _lib.ptr(dw), 0, m, _lib.ptr(ws), ws_bytes, _lib.stream()))
    return dw


class SparseConvFunction(Function):
    """forward / inverse / submanifold / 1x1 share one implementation: they differ only in which
    neighbour table feeds the forward (rulebook=None -> identity, the 1x1 shortcut; n_rows is then
    the live row count of the graph path, or None)."""

    @staticmethod
    def forward(ctx, features, weight, bias, rulebook, inverse, math, n_rows=None):
        feats = _f32c(features)
        kvol = 1 if rulebook is None else rulebook.kvol
        c_in, c_out = weight.shape[-2], weight.shape[-1]
        w3 = _f32c(weight).view(kvol, c_in, c_out)
        if rulebook is None:
            nbr, n_dst, n_src_dev, n_dst_dev = None, feats.shape[0], n_rows, n_rows
        elif inverse:
            nbr, n_dst = rulebook.nbr_in, rulebook.nbr_in.shape[0]
            n_src_dev, n_dst_dev = rulebook.n_out_dev, rulebook.n_in_dev
        else:
            nbr, n_dst = rulebook.nbr_out, rulebook.nbr_out.shape[0]
            n_src_dev, n_dst_dev = rulebook.n_in_dev, rulebook.n_out_dev
        b = _f32c(bias) if bias is not None else None
        out = conv_apply(feats, w3, 0, b, nbr, n_dst, c_out, math, n_src_dev, n_dst_dev)
        ctx.save_for_backward(feats, w3)
        ctx.rulebook, ctx.inverse, ctx.math = rulebook, inverse, math
        ctx.counts = (n_src_dev, n_dst_dev)
        ctx.has_bias = bias is not None
        ctx.w_shape, ctx.in_dtype = weight.shape, features.dtype
        return out.to(features.dtype) if features.dtype != torch.float32 else out

    @staticmethod
    def backward(ctx, grad_out):
        feats, w3 = ctx.saved_tensors
        rb, inverse, math = ctx.rulebook, ctx.inverse, ctx.math
        n_src_dev, n_dst_dev = ctx.counts
        g = _f32c(grad_out)
        kvol, c_in, c_out = w3.shape
        d_feats = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            if rb is None:
                nbr_t = None
            else:
                nbr_t = rb.nbr_out if inverse else rb.nbr_in
            d_feats = conv_apply(g, w3, 1, None, nbr_t, feats.shape[0], c_in, math, n_dst_dev, n_src_dev)
            if ctx.in_dtype != torch.float32:
                d_feats = d_feats.to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            if rb is None:
                pa = pb = pn = None
            else:
                pa, pb = (rb.pairs[1], rb.pairs[0]) if inverse else (rb.pairs[0], rb.pairs[1])
                pn = rb.pair_num
            d_w = conv_wgrad(feats, g, pa, pb, pn, kvol, math, n_src_dev, n_dst_dev).view(ctx.w_shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if n_dst_dev is None:
                d_b = g.sum(0)
            else:  # graph path: only the live rows of the capacity-sized gradient count
                d_b = live_col_sum(g, n_dst_dev)
        return d_feats, d_w, d_b, None, None, None, None


class ToDenseFunction(Function):
    @staticmethod
    def forward(ctx, features, indices, batch_size, h, w, n_rows=None):
        lib = _lib.load()
        _lib.require_cuda(features, indices)
        feats = _f32c(features)
        indices = indices.contiguous()
        n, c = feats.shape
        dense = torch.empty((batch_size, c, h, w), dtype=torch.float32, device=feats.device)
        table = torch.empty((max(batch_size * h * w, 1),), dtype=torch.int32, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.check(lib.wfsp_to_dense(_lib.ptr(feats), _lib.ptr(indices), n, _lib.ptr(n_rows), c, batch_size, h, w,
                                         _lib.ptr(dense), _lib.ptr(table), _lib.stream()))
        ctx.save_for_backward(indices)
        ctx.dims = (n, c, batch_size, h, w)
        ctx.n_rows = n_rows
        ctx.in_dtype = features.dtype
        return dense.to(features.dtype) if features.dtype != torch.float32 else dense

    @staticmethod
    def backward(ctx, grad_dense):
        lib = _lib.load()
        (indices,) = ctx.saved_tensors
        n, c, b, h, w = ctx.dims
        g = _f32c(grad_dense)
        d_feats = torch.empty((n, c), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(lib.wfsp_to_dense_bwd(_lib.ptr(g), _lib.ptr(indices), n, _lib.ptr(ctx.n_rows), c, b, h, w,
                                             _lib.ptr(d_feats), _lib.stream()))
        if ctx.in_dtype != torch.float32:
            d_feats = d_feats.to(ctx.in_dtype)
        return d_feats, None, None, None, None, None


class BatchNormReLUFunction(Function):
    """nn.BatchNorm1d (+ nn.ReLU) over the live rows of a capacity-sized feature buffer
    (wfsp_bn_relu_fwd / _bwd).  Used by the graph path only; the eager path keeps the stock
    torch modules the reference constructs (src/models/SPConvBlocks.py:505-508)."""

    @staticmethod
    def forward(ctx, x, n_rows, gamma, beta, running_mean, running_var, momentum, eps, training, relu):
        lib = _lib.load()
        _lib.require_cuda(x)
        x = _f32c(x)
        n, c = x.shape
        y = torch.empty_like(x)
        mean = torch.empty((c,), dtype=torch.float32, device=x.device)
        invstd = torch.empty((c,), dtype=torch.float32, device=x.device)
        ws_bytes = lib.wfsp_bn_workspace_bytes(n, c)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.wfsp_bn_relu_fwd(_lib.ptr(x), n, _lib.ptr(n_rows), c, _lib.ptr(gamma), _lib.ptr(beta),
                                            _lib.ptr(running_mean), _lib.ptr(running_var), float(momentum), float(eps),
                                            int(training), int(relu), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(invstd),
                                            _lib.ptr(ws), ws_bytes, _lib.stream()))
        ctx.save_for_backward(x, gamma, beta, mean, invstd)
        ctx.n_rows, ctx.relu = n_rows, relu
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, gamma, beta, mean, invstd = ctx.saved_tensors
        dy = _f32c(dy)
        n, c = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty((c,), dtype=torch.float32, device=x.device)
        db = torch.empty((c,), dtype=torch.float32, device=x.device)
        ws_bytes = lib.wfsp_bn_workspace_bytes(n, c)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.wfsp_bn_relu_bwd(_lib.ptr(x), _lib.ptr(dy), n, _lib.ptr(ctx.n_rows), c, _lib.ptr(gamma),
                                            _lib.ptr(beta), _lib.ptr(mean), _lib.ptr(invstd), int(ctx.relu),
                                            _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(ws), ws_bytes,
                                            _lib.stream()))
        return (dx, None, dg if gamma is not None else None, db if beta is not None else None, None, None, None, None,
                None, None)


def batch_norm_relu(x, n_rows, bn, relu):
    """Applies the nn.BatchNorm1d module `bn` (+ ReLU) to the live rows of x."""
    if bn.momentum is None:
        # torch's cumulative moving average needs 1 / num_batches_tracked on the host every step: not expressible
        # in a captured launch, and silently substituting 0.1 would change the running statistics
        raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative average) is not supported on the "
                                  "device-counted path; use a fixed momentum")
    momentum = bn.momentum
    training = bn.training or bn.running_mean is None
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return BatchNormReLUFunction.apply(x, n_rows, bn.weight, bn.bias, bn.running_mean, bn.running_var, momentum,
                                       bn.eps, training, relu)


def indice_conv(features, filters, indice_pairs_rulebook, bias=None, math="bf16"):
    return SparseConvFunction.apply(features, filters, bias, indice_pairs_rulebook, False, math, None)


def indice_inverse_conv(features, filters, rulebook, bias=None, math="bf16"):
    return SparseConvFunction.apply(features, filters, bias, rulebook, True, math, None)


indice_subm_conv = indice_conv
