"""Autograd functions over libwfsp.so (mirror of upstream spconv/functional.py:
indice_conv / indice_subm_conv / indice_inverse_conv and the dense scatter)."""
import torch
from torch.autograd import Function

from .. import _lib

_MATH = {"fp32": _lib.MATH_FP32, "bf16": _lib.MATH_BF16}


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def conv_apply(src, weight3, transpose_w, bias, nbr, n_dst, c_dst, math):
    """dst[r] = bias + sum_k src[nbr[r,k]] @ (weight3[k] or weight3[k]^T)   (wfsp_conv_apply)"""
    lib = _lib.load()
    _lib.require_cuda(src, weight3)
    kvol = weight3.shape[0]
    c_red = src.shape[1]
    dst = torch.empty((n_dst, c_dst), dtype=torch.float32, device=src.device)
    if n_dst == 0:
        return dst
    m = _MATH[math]
    ws_bytes = lib.wfsp_conv_apply_workspace_bytes(kvol, src.shape[0], c_red, c_dst, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=src.device) if ws_bytes else None
    with torch.cuda.device(src.device):
        _lib.check(lib.wfsp_conv_apply(_lib.ptr(src), src.shape[0], c_red, _lib.ptr(weight3), int(transpose_w),
                                       _lib.ptr(bias), _lib.ptr(nbr), kvol, _lib.ptr(dst), n_dst, c_dst, m,
                                       _lib.ptr(ws), ws_bytes, _lib.stream()))
    return dst


def conv_wgrad(a, b, pair_a, pair_b, pair_num, kvol, math):
    """d_weight[k] = sum_p a[pair_a[k,p]]^T (x) b[pair_b[k,p]]   (wfsp_conv_wgrad)"""
    lib = _lib.load()
    c_a, c_b = a.shape[1], b.shape[1]
    dw = torch.empty((kvol, c_a, c_b), dtype=torch.float32, device=a.device)
    pitch = pair_a.shape[-1] if pair_a is not None else a.shape[0]
    m = _MATH[math]
    ws_bytes = lib.wfsp_conv_wgrad_workspace_bytes(kvol, a.shape[0], c_a, b.shape[0], c_b, pitch, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=a.device) if ws_bytes else None
    with torch.cuda.device(a.device):
        _lib.check(lib.wfsp_conv_wgrad(_lib.ptr(a), a.shape[0], c_a, _lib.ptr(b), b.shape[0], c_b, _lib.ptr(pair_a),
                                       _lib.ptr(pair_b), _lib.ptr(pair_num), kvol, pitch, _lib.ptr(dw), 0, m,
                                       _lib.ptr(ws), ws_bytes, _lib.stream()))
    return dw


class SparseConvFunction(Function):
    """forward / inverse / submanifold / 1x1 share one implementation: they differ only in which
    neighbour table feeds the forward (rulebook=None -> identity, the 1x1 shortcut)."""

    @staticmethod
    def forward(ctx, features, weight, bias, rulebook, inverse, math):
        feats = _f32c(features)
        kvol = 1 if rulebook is None else rulebook.kvol
        c_in, c_out = weight.shape[-2], weight.shape[-1]
        w3 = _f32c(weight).view(kvol, c_in, c_out)
        if rulebook is None:
            nbr, n_dst = None, feats.shape[0]
        elif inverse:
            nbr, n_dst = rulebook.nbr_in, rulebook.nbr_in.shape[0]
        else:
            nbr, n_dst = rulebook.nbr_out, rulebook.nbr_out.shape[0]
        b = _f32c(bias) if bias is not None else None
        out = conv_apply(feats, w3, 0, b, nbr, n_dst, c_out, math)
        ctx.save_for_backward(feats, w3)
        ctx.rulebook, ctx.inverse, ctx.math = rulebook, inverse, math
        ctx.has_bias = bias is not None
        ctx.w_shape, ctx.in_dtype = weight.shape, features.dtype
        return out.to(features.dtype) if features.dtype != torch.float32 else out

    @staticmethod
    def backward(ctx, grad_out):
        feats, w3 = ctx.saved_tensors
        rb, inverse, math = ctx.rulebook, ctx.inverse, ctx.math
        g = _f32c(grad_out)
        kvol, c_in, c_out = w3.shape
        d_feats = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            if rb is None:
                nbr_t = None
            else:
                nbr_t = rb.nbr_out if inverse else rb.nbr_in
            d_feats = conv_apply(g, w3, 1, None, nbr_t, feats.shape[0], c_in, math)
            if ctx.in_dtype != torch.float32:
                d_feats = d_feats.to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            if rb is None:
                pa = pb = pn = None
            else:
                pa, pb = (rb.pairs[1], rb.pairs[0]) if inverse else (rb.pairs[0], rb.pairs[1])
                pn = rb.pair_num
            d_w = conv_wgrad(feats, g, pa, pb, pn, kvol, math).view(ctx.w_shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g.sum(0)
        return d_feats, d_w, d_b, None, None, None


class ToDenseFunction(Function):
    @staticmethod
    def forward(ctx, features, indices, batch_size, h, w):
        lib = _lib.load()
        _lib.require_cuda(features, indices)
        feats = _f32c(features)
        indices = indices.contiguous()
        n, c = feats.shape
        dense = torch.empty((batch_size, c, h, w), dtype=torch.float32, device=feats.device)
        table = torch.empty((max(batch_size * h * w, 1),), dtype=torch.int32, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.check(lib.wfsp_to_dense(_lib.ptr(feats), _lib.ptr(indices), n, c, batch_size, h, w,
                                         _lib.ptr(dense), _lib.ptr(table), _lib.stream()))
        ctx.save_for_backward(indices)
        ctx.dims = (n, c, batch_size, h, w)
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
            _lib.check(lib.wfsp_to_dense_bwd(_lib.ptr(g), _lib.ptr(indices), n, c, b, h, w, _lib.ptr(d_feats),
                                             _lib.stream()))
        if ctx.in_dtype != torch.float32:
            d_feats = d_feats.to(ctx.in_dtype)
        return d_feats, None, None, None, None


def indice_conv(features, filters, indice_pairs_rulebook, bias=None, math="bf16"):
    return SparseConvFunction.apply(features, filters, bias, indice_pairs_rulebook, False, math)


def indice_inverse_conv(features, filters, rulebook, bias=None, math="bf16"):
    return SparseConvFunction.apply(features, filters, bias, rulebook, True, math)


indice_subm_conv = indice_conv
