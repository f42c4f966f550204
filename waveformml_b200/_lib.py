"""ctypes binding of libwfsp.so (include/wfsp.h).  There is no fallback: if the library cannot be
loaded, importing the compute path raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwfsp.so")

F32, BF16, I16 = 0, 1, 2
MATH_FP32, MATH_BF16, MATH_BF16X3 = 0, 1, 2

_c = ctypes
_vp, _i64, _int, _sz, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_size_t, _c.c_float
_intp = _c.POINTER(_c.c_int)

# name -> (restype, argtypes); must list every symbol include/wfsp.h declares
SIGNATURES = {
    "wfsp_version": (_int, []),
    "wfsp_debug_trace": (_int, [_vp]),
    "wfsp_last_error": (_c.c_char_p, []),
    "wfsp_device_info": (_int, [_intp, _intp, _intp]),
    "wfsp_set_option": (_int, [_c.c_char_p, _int]),
    "wfsp_kernel_launches": (_c.c_ulonglong, []),
    "wfsp_source_hash": (_c.c_char_p, []),
    "wfsp_batch_pack": (_int, [_vp, _vp, _int, _i64, _vp, _int, _vp, _vp, _i64, _f32, _vp, _vp, _int, _i64, _vp]),
    "wfsp_conv_out_shape": (_int, [_intp] * 6),
    "wfsp_rulebook_workspace_bytes": (_sz, [_i64, _int, _intp, _intp]),
    "wfsp_rulebook_conv": (_int, [_vp, _i64, _vp, _int, _intp, _intp, _intp, _intp, _intp, _vp, _i64, _vp, _vp, _vp,
                                  _vp, _sz, _vp]),
    "wfsp_rulebook_subm": (_int, [_vp, _i64, _vp, _int, _intp, _intp, _intp, _vp, _vp, _vp, _sz, _vp]),
    "wfsp_rulebook_tables": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "wfsp_stage_inputs": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _c.c_int32, _vp]),
    "wfsp_rulebook_build": (_int, [_vp, _i64, _vp, _i64, _int, _intp, _intp, _intp, _intp, _intp, _int, _vp, _i64, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wfsp_conv_out_shape_nd": (_int, [_int, _intp, _intp, _intp, _intp, _intp, _intp]),
    "wfsp_rulebook_workspace_bytes_nd": (_sz, [_int, _i64, _int, _intp, _intp]),
    "wfsp_rulebook_build_nd": (_int, [_int, _vp, _i64, _vp, _i64, _int, _intp, _intp, _intp, _intp, _intp, _int, _vp,
                                      _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wfsp_rulebook_build_phased": (_int, [_int, _vp, _i64, _vp, _i64, _int, _intp, _intp, _intp, _intp, _intp, _int, _vp,
                                          _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _int, _intp, _vp]),
    "wfsp_conv_apply_workspace_bytes": (_sz, [_int, _i64, _int, _int, _int]),
    "wfsp_conv_apply": (_int, [_vp, _i64, _vp, _int, _vp, _int, _vp, _vp, _int, _vp, _i64, _vp, _i64, _int, _int, _vp,
                               _sz, _vp]),
    "wfsp_conv_wgrad_workspace_bytes": (_sz, [_int, _i64, _int, _i64, _int, _i64, _int]),
    "wfsp_conv_wgrad": (_int, [_vp, _i64, _vp, _int, _vp, _i64, _vp, _int, _vp, _vp, _vp, _int, _i64, _i64, _vp, _int,
                               _int, _vp, _sz, _vp]),
    "wfsp_to_dense": (_int, [_vp, _vp, _i64, _vp, _int, _int, _int, _int, _vp, _vp, _vp]),
    "wfsp_dense_cell_table": (_int, [_vp, _i64, _vp, _int, _int, _int, _vp, _vp]),
    "wfsp_to_dense_from_table": (_int, [_vp, _int, _int, _int, _int, _vp, _vp, _vp]),
    "wfsp_to_dense_bwd": (_int, [_vp, _vp, _i64, _vp, _int, _int, _int, _int, _vp, _vp]),
    "wfsp_bn_workspace_bytes": (_sz, [_i64, _int]),
    "wfsp_bn_relu_fwd": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _vp, _vp, _f32, _f32, _int, _int, _vp, _vp, _vp, _vp,
                                _sz, _vp]),
    "wfsp_bn_relu_bwd": (_int, [_vp, _vp, _i64, _vp, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    # (6) bf16-resident pipeline
    "wfsp_prepared_weight_bytes": (_sz, [_int, _int, _int]),
    "wfsp_prep_weights": (_int, [_vp, _int, _vp]),
    "wfsp_cast_rows_bf16": (_int, [_vp, _i64, _vp, _int, _vp, _vp]),
    "wfsp_conv_apply_bf16": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "wfsp_conv_apply_bf16_ex": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "wfsp_bn_relu_bwd_parts": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                                      _vp]),
    "wfsp_bn_partials_bytes": (_sz, [_i64, _int]),
    "wfsp_bn_relu_fwd_stats": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _int, _vp, _vp,
                                      _vp, _vp, _vp]),
    "wfsp_conv_wgrad_bf16": (_int, [_vp, _i64, _vp, _int, _vp, _i64, _vp, _int, _vp, _vp, _vp, _int, _i64, _i64, _vp,
                                    _int, _vp]),
    "wfsp_bn_relu_fwd_x": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _vp, _vp, _f32, _f32, _int, _int, _vp, _vp, _vp, _vp,
                                  _vp, _sz, _vp]),
    "wfsp_bn_relu_bwd_x": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                                  _sz, _vp]),
    "wfsp_dropout_factors": (_int, [_vp, _i64, _int, _vp, _vp]),
    "wfsp_bn_relu_fwd_stats_ex": (_int, [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _int, _vp, _vp,
                                         _vp, _vp, _vp, _vp]),
    "wfsp_bn_relu_bwd_x_ex": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                                     _sz, _vp, _vp]),
    "wfsp_act_fwd": (_int, [_vp, _i64, _vp, _int, _int, _vp, _vp, _vp]),
    "wfsp_act_bwd": (_int, [_vp, _vp, _i64, _vp, _int, _int, _vp, _vp, _vp]),
    "wfsp_sgd_step_p2p": (_int, [_vp, _vp, _vp, _i64, _f32, _f32, _int, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _int, _int,
                                 _int, _vp]),
    "wfsp_sgd_p2p_wait": (_int, [_vp, _vp, _int, _vp]),
    "wfsp_segment_l1_workspace_bytes": (_sz, [_i64]),
    "wfsp_segment_l1_fwd": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp]),
    "wfsp_segment_l1_bwd": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "wfsp_col_sum": (_int, [_vp, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "wfsp_head_workspace_bytes": (_sz, [_int, _int, _int]),
    "wfsp_head_tail_workspace_bytes": (_sz, [_int, _int, _int]),
    "wfsp_head_ce_tail": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "wfsp_head_ce_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _sz, _vp]),
    "wfsp_head_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wfsp_head_bwd_small": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "wfsp_window_edges_workspace_bytes": (_sz, [_i64]),
    "wfsp_window_edges": (_int, [_i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "wfsp_sgd_step": (_int, [_vp, _vp, _vp, _i64, _f32, _f32, _int, _f32, _f32, _vp]),
    "wfsp_sgd_step_ex": (_int, [_vp, _vp, _vp, _i64, _f32, _f32, _int, _f32, _f32, _int, _vp]),
}


class PrepJob(ctypes.Structure):
    """struct wfsp_prep_job (include/wfsp.h)"""
    _fields_ = [("weight", ctypes.c_void_p), ("out", ctypes.c_void_p), ("kvol", ctypes.c_int), ("c_red", ctypes.c_int),
                ("c_dst", ctypes.c_int), ("transpose_w", ctypes.c_int)]


class ConvEpilogue(ctypes.Structure):
    """struct wfsp_conv_epilogue (include/wfsp.h)"""
    _fields_ = [("bn_partials", ctypes.c_void_p), ("bwd_x", ctypes.c_void_p), ("bwd_mean", ctypes.c_void_p),
                ("bwd_invstd", ctypes.c_void_p), ("bwd_gamma", ctypes.c_void_p), ("bwd_beta", ctypes.c_void_p),
                ("bwd_partials", ctypes.c_void_p), ("bwd_relu", ctypes.c_int), ("k_split", ctypes.c_int),
                ("bn", ctypes.c_void_p)]


class BnFuse(ctypes.Structure):
    """struct wfsp_bn_fuse (include/wfsp.h)"""
    _fields_ = [("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("running_mean", ctypes.c_void_p),
                ("running_var", ctypes.c_void_p), ("momentum", ctypes.c_float), ("eps", ctypes.c_float),
                ("relu", ctypes.c_int), ("y", ctypes.c_void_p), ("y_bf16", ctypes.c_void_p),
                ("save_mean", ctypes.c_void_p), ("save_invstd", ctypes.c_void_p), ("dropout", ctypes.c_void_p),
                ("barrier", ctypes.c_void_p), ("n_rows_hint", ctypes.c_int64)]


class Dropout(ctypes.Structure):
    """struct wfsp_dropout (include/wfsp.h)"""
    _fields_ = [("p", ctypes.c_float), ("seed", ctypes.c_ulonglong), ("step_dev", ctypes.c_void_p), ("salt", ctypes.c_uint)]


def dropout_spec(p, seed, step_dev=None, salt=0):
    d = Dropout()
    d.p, d.seed, d.salt = float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(salt)
    d.step_dev = None if step_dev is None else step_dev.data_ptr()
    return d


def conv_epilogue(bn_partials=None, bwd=None, k_split=0, bn=None):
    """bwd = (x, mean, invstd, gamma, beta, relu, partials) tensors of the BatchNorm whose dy this dgrad produces.
    bn = a BnFuse (kept alive by the returned object): the BatchNorm behind this convolution."""
    def a(t):
        return None if t is None else t.data_ptr()
    e = ConvEpilogue()
    if bn is not None:
        e._bn_keep = bn
        e.bn = ctypes.addressof(bn)
    e.bn_partials = a(bn_partials)
    if bwd is not None:
        x, mean, invstd, gamma, beta, relu, partials = bwd
        e.bwd_x, e.bwd_mean, e.bwd_invstd, e.bwd_gamma, e.bwd_beta = a(x), a(mean), a(invstd), a(gamma), a(beta)
        e.bwd_relu, e.bwd_partials = int(relu), a(partials)
    e.k_split = int(k_split)
    return e


_lib = None


class WfspError(RuntimeError):
    pass


EXPECTED_VERSION = 203  # include/wfsp.h WFSP_VERSION: bumped with every change of the C ABI


def load():
    """Loads libwfsp.so.  The library must have been compiled from the sources of THIS tree (digest check,
    build.source_hash): a missing or stale binary is rebuilt with nvcc first, and if that is impossible the load
    fails -- a stale .so called through new ctypes signatures would corrupt memory silently."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build._stale():
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.wfsp_version() != EXPECTED_VERSION:
        raise WfspError("libwfsp.so has ABI version %d, the bindings expect %d" % (lib.wfsp_version(), EXPECTED_VERSION))
    if lib.wfsp_source_hash().decode() != _build.source_hash():
        raise WfspError("libwfsp.so was built from other sources (%s, tree is %s)"
                        % (lib.wfsp_source_hash().decode(), _build.source_hash()))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise WfspError("libwfsp error %d: %s" % (rc, load().wfsp_last_error().decode(errors="replace")))


def ints(v, n=2):
    try:
        v = [int(x) for x in v]
    except TypeError:
        v = [int(v)] * n
    if len(v) != n:
        raise ValueError("expected %d values, got %r" % (n, v))
    return (ctypes.c_int * n)(*v)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("waveformml_b200 has no CPU path: tensor on %s (the sparse-conv kernels are CUDA "
                               "sm_100a only)" % t.device)
