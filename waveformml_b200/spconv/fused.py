"""bf16-resident execution of a whole SparseSequential stack (include/wfsp.h section 6).

The reference builds its sparse networks as spconv.SparseSequential(SparseConv2d, BatchNorm1d, ReLU,
..., ToDense) (src/models/SPConvBlocks.py:498-516, :261-313, :756-822).  Run module by module, every
layer boundary costs a cast of the activations to the tensor-core operand type (once in forward, again
in wgrad, and the gradient twice: dgrad + wgrad) and a weight re-layout per call.  Here the stack is
one autograd node:

  forward   one launch prepares the bf16 weights of every layer (forward + dgrad layouts);
            conv (bf16 in, fp32 out) -> BatchNorm statistics -> normalise + ReLU written directly as the
            next layer's bf16 operand; only the last block writes fp32 (what ToDense / the caller reads);
  backward  BatchNorm/ReLU backward writes the bf16 gradient that wgrad and dgrad both read;
            dgrad's fp32 output is the previous block's dy.

Arithmetic is unchanged: the same values are rounded to bf16 at the same points as on the per-layer
path (conv operands only), accumulation and BatchNorm stay fp32.  Anything the plan does not cover
(fp32 math mode, eval-mode BatchNorm under autograd, Dropout in training, forward hooks on inner
modules, unknown modules) falls back to the per-layer path in SparseSequential.forward.
"""
import ctypes
import os

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib
from . import functional as Fsp

_enabled = True
# BatchNorm statistics come from the convolution epilogue (per-32-row partials) whenever the output has more
# rows than this; small outputs then normalise in ONE launch that folds the few partials itself
_STATS_FUSE_MIN_ROWS = 0
# The dgrad epilogue can also take the two reductions of the previous block's BatchNorm backward (sum dy', sum dy' xhat;
# wfsp_conv_epilogue.bwd_partials + wfsp_bn_relu_bwd_parts).  Measured on B200 and left OFF: the epilogue of these
# kernels is not overlapped with the next tile's main loop (one CTA per SM), so the extra pass over x is exposed --
# C5@1024: 3x3 dgrad 156 -> 451 us against 141 us of bn_bwd_partial saved; 64 events: dgrad +9 us against -2.5 us.
_BWD_PARTS = False
# Graph-topology choices of the fused stack, each measured on the 64-event step (WFSP_TOPOLOGY=0 switches them off for
# an A/B run): ToDense's cell table built on the geometry branch; layer 0's weights prepared in a launch of their own.
_EARLY_CELL_TABLE = os.environ.get("WFSP_TOPOLOGY", "1") != "0"
_SPLIT_PREP = os.environ.get("WFSP_TOPOLOGY", "1") != "0"


def set_fused(flag):
    """Globally enable / disable the fused stack path (per-layer execution is always available)."""
    global _enabled
    _enabled = bool(flag)


def is_enabled():
    return _enabled


def pitch8(c):
    return (c + 7) // 8 * 8


def is_operand_format(features, channels):
    """bf16 [N, channels rounded up to 8], contiguous: the layout the tensor-core kernels gather from."""
    return (features.dtype == torch.bfloat16 and features.dim() == 2 and features.shape[1] == pitch8(channels)
            and features.is_contiguous())


class Block:
    __slots__ = ("conv", "bn", "relu", "drop")

    def __init__(self, conv):
        self.conv, self.bn, self.relu, self.drop = conv, None, False, 0.0


class Plan:
    def __init__(self, blocks, to_dense):
        self.blocks, self.to_dense = blocks, to_dense
        self.prepared = None  # weight preparation already launched for this step (prepare_stacks)

    def params(self):
        out = []
        for b in self.blocks:
            out.append(b.conv.weight)
            out.append(b.conv.bias)
            out.append(b.bn.weight if b.bn is not None else None)
            out.append(b.bn.bias if b.bn is not None else None)
        return out


def compile_stack(modules, sparse_conv_cls, to_dense_cls):
    """modules: list of the SparseSequential's children.  Returns a Plan or None if the stack contains
    something the fused path does not cover."""
    blocks, to_dense, cur = [], False, None
    n = len(modules)
    for i, m in enumerate(modules):
        if getattr(m, "_forward_hooks", None) or getattr(m, "_forward_pre_hooks", None):
            return None  # someone is observing per-layer activations (e.g. scripts/PlotModelWeights.py:44-53)
        if isinstance(m, sparse_conv_cls):
            cur = Block(m)
            blocks.append(cur)
        elif isinstance(m, nn.BatchNorm1d):
            if cur is None or cur.bn is not None or cur.relu or m.momentum is None or not m.track_running_stats:
                return None
            if not m.training and torch.is_grad_enabled():
                return None  # eval-mode BN backward is affine, not the batch-statistics formula
            cur.bn = m
            if getattr(m, "fused_relu", False):  # sparseconvnet.BatchNormReLU facade
                cur.relu = True
        elif isinstance(m, nn.ReLU):
            if cur is None or cur.relu:
                return None
            cur.relu = True
        elif isinstance(m, nn.Identity) or (isinstance(m, nn.Dropout) and (not m.training or m.p == 0)):
            continue
        elif isinstance(m, nn.Dropout):
            # training-mode Dropout behind a block's BatchNorm(+ReLU) (SPConvBlocks.py:375-376, 509-510): fused into
            # the kernel that writes the block's output; the backward kernels regenerate the mask
            if (cur is None or cur.bn is None or cur.drop or not cur.bn.training or not (0.0 < m.p < 1.0)
                    or cur.conv.out_channels > 512):
                return None
            cur.drop = float(m.p)
        elif isinstance(m, to_dense_cls):
            if i != n - 1 or cur is None:
                return None
            to_dense = True
        else:
            return None
    if not blocks:
        return None
    return Plan(blocks, to_dense)


_write_through = 0
_write_through_zeroed = False  # the write-through targets were zeroed before this backward (grad_write_through(zeroed=True))


class grad_write_through:
    """Context manager: inside it, the fused kernels may OVERWRITE `param._wfsp_grad_out` with the gradient and hand
    None to autograd.  That is only right when the step runs exactly one backward over freshly zeroed gradients and
    no parameter is shared between two fused nodes (harness.TrainStep.forward_backward) -- so it is opt-in per step;
    everywhere else (gradient accumulation over micro-batches, retain_graph, shared parameters) the kernels return
    fresh tensors and torch.autograd accumulates as usual."""

    def __init__(self, zeroed=False):
        self.zeroed = zeroed

    def __enter__(self):
        global _write_through, _write_through_zeroed
        _write_through += 1
        self._prev = _write_through_zeroed
        _write_through_zeroed = bool(self.zeroed)
        return self

    def __exit__(self, *exc):
        global _write_through, _write_through_zeroed
        _write_through -= 1
        _write_through_zeroed = self._prev
        return False


def _grad_target(param, shape, dev):
    """Where a parameter gradient is written.  A training harness that owns a flat gradient buffer and
    runs ONE backward per step can attach `param._wfsp_grad_out` (a contiguous fp32 view of the parameter's
    shape, e.g. harness.FlatGrads) and run the step inside `grad_write_through()`: the kernel then writes the
    gradient there and autograd gets None -- no accumulation kernel per parameter.  Otherwise a fresh tensor is
    returned to autograd as usual."""
    tgt = getattr(param, "_wfsp_grad_out", None) if (param is not None and _write_through > 0) else None
    if tgt is not None and tgt.dtype == torch.float32 and tgt.is_contiguous() and tgt.numel() == param.numel() \
            and tgt.device == dev:
        return tgt.view(shape), True
    return torch.empty(shape, dtype=torch.float32, device=dev), False


# Data-parallel training: a harness that exchanges gradients in buckets sets `grads_ready_hook(params, events)`; it is
# called during backward as soon as the gradients of `params` are final, with the events (on the main / side stream)
# after which they may be read -- so the exchange of the later layers overlaps the backward pass of the earlier ones
# (the reference's DDP does the same with its gradient buckets, src/utils/util.py:233-236).
grads_ready_hook = None

_side_streams = {}


def _side_stream(dev, which=0):
    """One extra stream per device.  The geometry of a stack (rulebooks: indices only) does not depend on
    the features, and wgrad does not feed the dgrad chain, so both run beside the main stream and join it
    through events -- inside a CUDA graph capture these become parallel branches of the graph."""
    key = (dev.type, dev.index, which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def _bn_ws(lib, n, c, dev):
    return torch.empty((lib.wfsp_bn_workspace_bytes(n, c),), dtype=torch.uint8, device=dev)


def _dense_geometry(t, idx=None):
    """(H, W, index rows [n, 3]) for the ToDense kernels; a 3-d tensor is scattered as [B, C, H, W*T] with its last
    two coordinates merged (the memory layout of [B, C, H, W, T])."""
    if len(t.spatial_shape) == 2:
        return t.spatial_shape[0], t.spatial_shape[1], (t.indices.contiguous() if idx is None else idx)
    h, w, d = t.spatial_shape
    if idx is None:
        i = t.indices
        idx = torch.stack([i[:, 0], i[:, 1], i[:, 2] * d + i[:, 3]], dim=1).contiguous()
    return h, w * d, idx


_drop_steps = {}
last_drop_seed = 0
_bn_barriers = {}


def _bn_barrier(dev):
    """Per-device counter pair of the grid barrier inside convolution launches that finish their BatchNorm themselves
    (zero once, left consistent by every launch)."""
    key = (dev.type, dev.index)
    if key not in _bn_barriers:
        _bn_barriers[key] = torch.zeros((2,), dtype=torch.int32, device=dev)
    return _bn_barriers[key]


def _ptr_int(t):
    return None if t is None else t.data_ptr()


def _drop_step(dev):
    """Per-device int64 counter mixed into every Dropout mask: advanced once per forward of a stack with Dropout (on
    the device, so a replayed CUDA graph draws a fresh mask every step)."""
    key = (dev.type, dev.index)
    if key not in _drop_steps:
        _drop_steps[key] = torch.zeros((1,), dtype=torch.int64, device=dev)
    return _drop_steps[key]


def _param_key(params):
    return tuple(0 if p is None else p.data_ptr() for p in params[0::4])


def _prep_weights(plan, params, need_in_grad, dev, fork):
    """Launches the preparation of every layer's weights (forward + dgrad layouts, one launch) on the second side
    stream, ordered after `fork`; the BatchNorm step counters ride along.  Returns the buffers and events."""
    lib = _lib.load()
    blocks = plan.blocks
    st = _lib.stream
    jobs, offs, total = [], [], 0
    for bi, b in enumerate(blocks):
        kvol = 1 if b.conv.conv1x1 else int(np.prod(b.conv.kernel_size))
        cin, cout = b.conv.in_channels, b.conv.out_channels
        f_off = total
        total += lib.wfsp_prepared_weight_bytes(kvol, cin, cout)
        d_off = None
        if bi > 0 or need_in_grad:
            d_off = total
            total += lib.wfsp_prepared_weight_bytes(kvol, cout, cin)
        offs.append((f_off, d_off))
    wbuf = torch.empty((total,), dtype=torch.uint8, device=dev)
    for bi, b in enumerate(blocks):
        kvol = 1 if b.conv.conv1x1 else int(np.prod(b.conv.kernel_size))
        cin, cout = b.conv.in_channels, b.conv.out_channels
        w = params[4 * bi]
        assert w.dtype == torch.float32 and w.is_contiguous()
        f_off, d_off = offs[bi]
        jobs.append(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr() + f_off, kvol, cin, cout, 0))
        if d_off is not None:
            jobs.append(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr() + d_off, kvol, cout, cin, 1))
    arr = (_lib.PrepJob * len(jobs))(*jobs)
    # Layer 0's forward layout: a short launch of its own on the CALLING stream (at the head of the step's main chain in
    # harness.GraphTrainStep), so the first convolution has no cross-branch join in front of it -- roots of parallel
    # graph branches start staggered by 1-3 us each and every join costs 2-4 us more.  Everything else is prepared on
    # a side stream, beside the batcher / the input cast / the first layer.
    split = _SPLIT_PREP and len(jobs) > 1
    w0_ready = None
    if split:
        if any(b.drop for b in blocks):  # the first BatchNorm kernel must see this step's value
            _drop_step(dev).add_(1)
        _lib.check(lib.wfsp_prep_weights(ctypes.cast(arr, ctypes.c_void_p), 1, st()))
        w0_ready = torch.cuda.Event()
        w0_ready.record(torch.cuda.current_stream())
    side2 = _side_stream(dev, 1)
    with torch.cuda.stream(side2):
        # (split: behind layer 0's launch, which then is the step's only root -- see harness.GraphTrainStep._body)
        side2.wait_event(w0_ready if split else fork)
        if split:
            rest = (_lib.PrepJob * (len(jobs) - 1))(*jobs[1:])
            _lib.check(lib.wfsp_prep_weights(ctypes.cast(rest, ctypes.c_void_p), len(jobs) - 1, st()))
        else:
            if any(b.drop for b in blocks):  # before w_ready: the first BatchNorm kernel must see this step's value
                _drop_step(dev).add_(1)
            _lib.check(lib.wfsp_prep_weights(ctypes.cast(arr, ctypes.c_void_p), len(jobs), st()))
        w_ready = torch.cuda.Event()
        w_ready.record(side2)
        if w0_ready is None:
            w0_ready = w_ready
        # BatchNorm step counters: one multi-tensor add here instead of one launch per layer on the main stream
        counters = [b.bn.num_batches_tracked for b in blocks
                    if b.bn is not None and b.bn.training and b.bn.num_batches_tracked is not None]
        if counters:
            torch._foreach_add_(counters, 1)
        side2_done = torch.cuda.Event()
        side2_done.record(side2)
    return {"wbuf": wbuf, "offs": offs, "w_ready": w_ready, "w0_ready": w0_ready, "done": side2_done, "need_in_grad": need_in_grad,
            "key": _param_key(params)}


def prepare_stacks(model):
    """Called at the very start of a training step (harness.GraphTrainStep._body): starts the weight preparation of
    every fused sparse stack of `model` now, beside the batcher, instead of when the stack's forward is reached.  The
    next forward of each stack picks the result up (and falls back to preparing itself if anything changed)."""
    from . import SparseSequential, SparseConvolution, ToDense, get_math_mode
    if not (_enabled and torch.is_grad_enabled()):
        return
    for seq in model.modules():
        if not isinstance(seq, SparseSequential):
            continue
        mods = list(seq._modules.values())
        if not all((m.math or get_math_mode()) == "bf16" for m in mods if isinstance(m, SparseConvolution)):
            continue
        plan = compile_stack(mods, SparseConvolution, ToDense)
        if plan is None:
            continue
        params = plan.params()
        w0 = params[0]
        if not w0.is_cuda:
            continue
        with torch.cuda.device(w0.device):
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream())
            seq._wfsp_prepared = _prep_weights(plan, params, False, w0.device, fork)


class FusedStackFunction(Function):
    @staticmethod
    def forward(ctx, plan, x, holder, features, *params):
        lib = _lib.load()
        _lib.require_cuda(features, x.indices)
        dev = features.device
        blocks = plan.blocks
        # bf16 features in the operand format ([N, C rounded up to 8], e.g. straight from batcher.pack_batch) are
        # used as they are; anything else is brought to fp32 and cast once
        ready16 = is_operand_format(features, blocks[0].conv.in_channels)
        feats = features if (ready16 or features.dtype == torch.float32) else features.float()
        feats = feats.contiguous()
        st = _lib.stream
        need_in_grad = features.requires_grad and not ready16
        with torch.cuda.device(dev):
            # ---- every layer's weights -> bf16 tensor-core layouts, one launch (already under way if the caller
            # announced the step early: prepare_stacks)
            main, side = torch.cuda.current_stream(), _side_stream(dev)
            fork = torch.cuda.Event()
            fork.record(main)
            prep = plan.prepared
            if prep is None or prep["need_in_grad"] != need_in_grad or prep["key"] != _param_key(params):
                if prep is not None:
                    main.wait_event(prep["done"])  # an early preparation that does not fit this call: just join it
                prep = _prep_weights(plan, params, need_in_grad, dev, fork)
            wbuf, offs, w_ready, side2_done = prep["wbuf"], prep["offs"], prep["w_ready"], prep["done"]
            ready_ev = getattr(features, "_wfsp_ready", None)
            if ready_ev is not None:  # features written on another stream (harness: waveform conversion beside the indices)
                main.wait_event(ready_ev)

            # ---- input activations -> bf16 once
            n0, c0 = feats.shape
            if ready16:
                a16 = feats
            else:
                a16 = torch.empty((max(n0, 1), pitch8(c0)), dtype=torch.bfloat16, device=dev)
                if n0:
                    _lib.check(lib.wfsp_cast_rows_bf16(_lib.ptr(feats), n0, _lib.ptr(x.n_rows), c0, _lib.ptr(a16), st()))

            # ---- geometry of every block on the side stream (overlaps weight preparation and the first layers)
            geoms, gcur = [], x
            with torch.cuda.stream(side):
                side.wait_event(fork)
                # graph path: only what the forward pass waits for (output rows, nbr_out) is built here; the pair lists
                # and nbr_in that backward needs follow on a third stream, beside the forward pass (an inverse
                # convolution reads nbr_in in its FORWARD pass: such stacks build whole rulebooks)
                split_rb = not any(b.conv.inverse for b in blocks)
                for b in blocks:
                    rb, outids, out_shape, out_rows = b.conv.geometry(gcur, front_only=split_rb)
                    ready = torch.cuda.Event()
                    ready.record(side)
                    nxt = x.__class__(None, outids, out_shape, gcur.batch_size, n_rows=out_rows)
                    nxt.indice_dict, nxt.grid = gcur.indice_dict, gcur.grid
                    geoms.append((rb, outids, out_shape, out_rows, gcur, nxt, ready))
                    gcur = nxt
                # ToDense: the cell table (row of every dense cell) depends on the output coordinates only -- built
                # here, beside the convolutions, so that only the scatter itself follows the last layer
                dense_pre = None
                if plan.to_dense and _EARLY_CELL_TABLE:
                    dh, dw, didx = _dense_geometry(gcur)
                    dtable = torch.empty((max(gcur.batch_size * dh * dw, 1),), dtype=torch.int32, device=dev)
                    _lib.check(lib.wfsp_dense_cell_table(_lib.ptr(didx), didx.shape[0], _lib.ptr(gcur.n_rows),
                                                         gcur.batch_size, dh, dw, _lib.ptr(dtable), st()))
                    dready = torch.cuda.Event()
                    dready.record(side)
                    dense_pre = (dh, dw, didx, dtable, dready)

            backs = [g[0] for g in geoms if g[0] is not None and getattr(g[0], "_pending", None)]
            back_done = None
            if backs:
                side3 = _side_stream(dev, 3)
                with torch.cuda.stream(side3):
                    side3.wait_event(geoms[-1][6])
                    for rb in backs:
                        rb.finish()
                    back_done = torch.cuda.Event()
                    back_done.record(side3)
            main.wait_event(prep["w0_ready"])  # layer 0's weights; the rest is waited for in front of layer 1
            drop_seed = 0
            if any(b.drop for b in blocks):  # from torch's CPU generator: reproducible under torch.manual_seed
                drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            saved, cur, out32 = [], x, None
            for bi, b in enumerate(blocks):
                conv = b.conv
                rb, outids, out_shape, out_rows, cur, nxt, ready = geoms[bi]
                main.wait_event(ready)
                if bi == 1 and prep["w0_ready"] is not w_ready:
                    main.wait_event(w_ready)
                last = bi == len(blocks) - 1
                kvol = 1 if rb is None else rb.kvol
                cin, cout = conv.in_channels, conv.out_channels
                if rb is None:
                    nbr, n_dst, n_src_dev, n_dst_dev = None, cur.indices.shape[0], cur.n_rows, cur.n_rows
                elif conv.inverse:
                    nbr, n_dst = rb.nbr_in, rb.nbr_in.shape[0]
                    n_src_dev, n_dst_dev = rb.n_out_dev, rb.n_in_dev
                else:
                    nbr, n_dst = rb.nbr_out, rb.nbr_out.shape[0]
                    n_src_dev, n_dst_dev = rb.n_in_dev, rb.n_out_dev
                n_src = a16.shape[0] if cur.indices.shape[0] else 0
                xf = torch.empty((n_dst, cout), dtype=torch.float32, device=dev)
                bias = params[4 * bi + 1]
                # large outputs followed by a training-mode BatchNorm: the conv epilogue also emits the
                # per-32-row-chunk statistics, so BatchNorm does not re-read the output to get them
                partials = None
                if b.bn is not None and b.bn.training and n_dst > _STATS_FUSE_MIN_ROWS and cout <= 512:
                    partials = torch.empty((lib.wfsp_bn_partials_bytes(n_dst, cout),), dtype=torch.uint8, device=dev)
                hint = Fsp.hints.get(n_dst_dev) if n_dst else 0
                # ---- BatchNorm / ReLU (/ Dropout) -> next operand (bf16) or the stack's output (fp32).  With the
                # epilogue statistics the BatchNorm is handed to the convolution launch itself (wfsp_bn_fuse): small
                # launches finish it behind a grid barrier, otherwise the library runs it right behind the convolution
                y32 = y16 = mean = invstd = None
                bn_in_conv = None
                if b.bn is not None or b.relu:
                    if last:
                        y32 = torch.empty_like(xf)
                    else:
                        y16 = torch.empty((max(n_dst, 1), pitch8(cout)), dtype=torch.bfloat16, device=dev)
                if b.bn is not None:
                    bn = b.bn
                    mean = torch.empty((cout,), dtype=torch.float32, device=dev)
                    invstd = torch.empty((cout,), dtype=torch.float32, device=dev)
                    if n_dst and partials is not None:
                        spec = _lib.dropout_spec(b.drop, drop_seed, _drop_step(dev), bi) if b.drop else None
                        bn_in_conv = _lib.BnFuse()
                        bn_in_conv.gamma, bn_in_conv.beta = _ptr_int(bn.weight), _ptr_int(bn.bias)
                        bn_in_conv.running_mean, bn_in_conv.running_var = _ptr_int(bn.running_mean), _ptr_int(bn.running_var)
                        bn_in_conv.momentum, bn_in_conv.eps, bn_in_conv.relu = float(bn.momentum), float(bn.eps), int(b.relu)
                        bn_in_conv.y, bn_in_conv.y_bf16 = _ptr_int(y32), _ptr_int(y16)
                        bn_in_conv.save_mean, bn_in_conv.save_invstd = _ptr_int(mean), _ptr_int(invstd)
                        bn_in_conv._spec_keep = spec
                        bn_in_conv.dropout = None if spec is None else ctypes.addressof(spec)
                        bn_in_conv.barrier = _ptr_int(_bn_barrier(dev))
                        bn_in_conv.n_rows_hint = hint
                if n_dst:
                    ep = _lib.conv_epilogue(bn_partials=partials, bn=bn_in_conv)
                    _lib.check(lib.wfsp_conv_apply_bf16_ex(_lib.ptr(a16), cur.indices.shape[0], _lib.ptr(n_src_dev), cin,
                                                           ctypes.c_void_p(wbuf.data_ptr() + offs[bi][0]), _lib.ptr(bias),
                                                           _lib.ptr(nbr), kvol, _lib.ptr(xf), n_dst, _lib.ptr(n_dst_dev),
                                                           hint, cout, ctypes.byref(ep), st()))
                if b.bn is not None or b.relu:
                    if b.bn is not None:
                        if bn_in_conv is not None:
                            pass  # done by (or right behind) the convolution launch
                        elif n_dst:
                            assert not b.drop, "Dropout needs the epilogue statistics path"
                            ws = _bn_ws(lib, n_dst, cout, dev)
                            _lib.check(lib.wfsp_bn_relu_fwd_x(
                                _lib.ptr(xf), n_dst, _lib.ptr(n_dst_dev), cout, _lib.ptr(bn.weight), _lib.ptr(bn.bias),
                                _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), float(bn.momentum), float(bn.eps),
                                int(bn.training), int(b.relu), _lib.ptr(y32), _lib.ptr(y16), _lib.ptr(mean),
                                _lib.ptr(invstd), _lib.ptr(ws), ws.numel(), st()))
                    elif n_dst:
                        _lib.check(lib.wfsp_act_fwd(_lib.ptr(xf), n_dst, _lib.ptr(n_dst_dev), cout, 1, _lib.ptr(y32),
                                                    _lib.ptr(y16), st()))
                elif last:
                    y32 = xf
                else:
                    y16 = torch.empty((max(n_dst, 1), pitch8(cout)), dtype=torch.bfloat16, device=dev)
                    if n_dst:
                        _lib.check(lib.wfsp_cast_rows_bf16(_lib.ptr(xf), n_dst, _lib.ptr(n_dst_dev), cout,
                                                           _lib.ptr(y16), st()))
                keep_x = xf if (b.bn is not None or b.relu) else None
                saved.append((a16, keep_x, mean, invstd, rb, cur.indices.shape[0], n_dst, n_src_dev, n_dst_dev, hint))
                cur, a16, out32 = nxt, y16, y32

            main.wait_event(side2_done)
            if back_done is not None:
                main.wait_event(back_done)
            holder["tensor"] = cur  # geometry of the stack's output
            ctx.plan, ctx.saved, ctx.offs, ctx.wbuf = plan, saved, offs, wbuf
            ctx.drop_seed = drop_seed
            global last_drop_seed
            last_drop_seed = drop_seed  # (tests: rebuild the masks of this forward with wfsp_dropout_factors)
            ctx.params = params
            ctx.need_in_grad, ctx.in_dtype = need_in_grad, features.dtype
            ctx.final = cur
            if plan.to_dense:
                n, c = out32.shape
                if dense_pre is not None:
                    h, w, ctx.dense_idx, table, dready = dense_pre
                    dense = torch.empty((cur.batch_size, c, h, w), dtype=torch.float32, device=dev)
                    main.wait_event(dready)
                    _lib.check(lib.wfsp_to_dense_from_table(_lib.ptr(out32), c, cur.batch_size, h, w, _lib.ptr(table),
                                                            _lib.ptr(dense), st()))
                    ctx.dense_table = table  # allocated on the side stream: kept until backward has run
                else:
                    h, w, ctx.dense_idx = _dense_geometry(cur)
                    dense = torch.empty((cur.batch_size, c, h, w), dtype=torch.float32, device=dev)
                    table = torch.empty((max(cur.batch_size * h * w, 1),), dtype=torch.int32, device=dev)
                    _lib.check(lib.wfsp_to_dense(_lib.ptr(out32), _lib.ptr(ctx.dense_idx), n, _lib.ptr(cur.n_rows),
                                                 c, cur.batch_size, h, w, _lib.ptr(dense), _lib.ptr(table), st()))
                return dense.view(cur.batch_size, c, *cur.spatial_shape)
            return out32

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        plan, saved, offs, wbuf, params = ctx.plan, ctx.saved, ctx.offs, ctx.wbuf, ctx.params
        blocks = plan.blocks
        dev = grad_out.device
        st = _lib.stream
        grads = [None] * len(params)
        with torch.cuda.device(dev):
            g = grad_out if grad_out.dtype == torch.float32 else grad_out.float()
            g = g.contiguous()
            final = ctx.final
            if plan.to_dense:
                n, c = saved[-1][6], blocks[-1].conv.out_channels
                h, w, _ = _dense_geometry(final, ctx.dense_idx)
                dy = torch.empty((n, c), dtype=torch.float32, device=dev)
                _lib.check(lib.wfsp_to_dense_bwd(_lib.ptr(g), _lib.ptr(ctx.dense_idx), n,
                                                 _lib.ptr(final.n_rows), c, final.batch_size, h, w, _lib.ptr(dy), st()))
            else:
                dy = g
            main, side = torch.cuda.current_stream(), _side_stream(dev)
            side_used = False
            # Tensors the side stream reads (the bf16 gradient of every block) must outlive the loop: the caching
            # allocator only knows the stream a block was ALLOCATED on, so a g16 dropped here could be handed to the
            # next iteration's main-stream buffers while wgrad on `side` still reads it (inside a capture the same
            # aliasing would be baked into the graph with no edge between the branches).  They are released after
            # main has waited for the side stream.
            cross_stream = []
            dy_parts = None  # BatchNorm-backward partial sums of `dy`, when the dgrad that produced it took them
            for bi in range(len(blocks) - 1, -1, -1):
                b = blocks[bi]
                conv = b.conv
                a16, xf, mean, invstd, rb, n_in, n_dst, n_src_dev, n_dst_dev, dst_hint = saved[bi]
                cin, cout = conv.in_channels, conv.out_channels
                kvol = 1 if rb is None else rb.kvol
                w_p, bias_p, gamma_p, beta_p = params[4 * bi: 4 * bi + 4]
                want_bias = bias_p is not None and ctx.needs_input_grad[4 + 4 * bi + 1]
                g16 = torch.empty((max(n_dst, 1), pitch8(cout)), dtype=torch.bfloat16, device=dev)
                cross_stream.append(g16)
                dx32 = torch.empty((n_dst, cout), dtype=torch.float32, device=dev) if want_bias else None
                if b.bn is not None:
                    dgam, gam_through = _grad_target(gamma_p, (cout,), dev)
                    dbet, bet_through = _grad_target(beta_p, (cout,), dev)
                    if dy_parts is not None:
                        _lib.check(lib.wfsp_bn_relu_bwd_parts(
                            _lib.ptr(xf), _lib.ptr(dy), n_dst, _lib.ptr(n_dst_dev), dst_hint, cout, _lib.ptr(gamma_p),
                            _lib.ptr(beta_p), _lib.ptr(mean), _lib.ptr(invstd), int(b.relu), _lib.ptr(dy_parts),
                            _lib.ptr(dx32), _lib.ptr(g16), _lib.ptr(dgam), _lib.ptr(dbet), st()))
                    else:
                        ws = _bn_ws(lib, max(n_dst, 1), cout, dev)
                        spec = _lib.dropout_spec(b.drop, ctx.drop_seed, _drop_step(dev), bi) if b.drop else None
                        _lib.check(lib.wfsp_bn_relu_bwd_x_ex(
                            _lib.ptr(xf), _lib.ptr(dy), n_dst, _lib.ptr(n_dst_dev), dst_hint, cout, _lib.ptr(gamma_p), _lib.ptr(beta_p),
                            _lib.ptr(mean), _lib.ptr(invstd), int(b.relu), _lib.ptr(dx32), _lib.ptr(g16), _lib.ptr(dgam),
                            _lib.ptr(dbet), _lib.ptr(ws), ws.numel(), None if spec is None else ctypes.byref(spec), st()))
                    if gamma_p is not None and not gam_through:
                        grads[4 * bi + 2] = dgam
                    if beta_p is not None and not bet_through:
                        grads[4 * bi + 3] = dbet
                elif n_dst:
                    if b.relu:
                        _lib.check(lib.wfsp_act_bwd(_lib.ptr(xf), _lib.ptr(dy), n_dst, _lib.ptr(n_dst_dev), cout, 1,
                                                    _lib.ptr(dx32), _lib.ptr(g16), st()))
                    else:
                        _lib.check(lib.wfsp_cast_rows_bf16(_lib.ptr(dy), n_dst, _lib.ptr(n_dst_dev), cout, _lib.ptr(g16),
                                                           st()))
                        dx32 = dy if want_bias else None
                if want_bias:
                    if n_dst_dev is None:
                        grads[4 * bi + 1] = dx32.sum(0)
                    else:  # graph path: only the live rows of the capacity-sized gradient count
                        grads[4 * bi + 1] = Fsp.live_col_sum(dx32, n_dst_dev)
                # ---- wgrad
                if ctx.needs_input_grad[4 + 4 * bi]:
                    dw, w_through = _grad_target(w_p, (kvol, cin, cout), dev)
                    if rb is None:
                        pa = pb = pn = None
                        pitch = n_in
                    else:
                        pa, pb = (rb.pairs[1], rb.pairs[0]) if conv.inverse else (rb.pairs[0], rb.pairs[1])
                        pn, pitch = rb.pair_num, rb.pairs.shape[-1]
                    hint = 0
                    if n_src_dev is not None:
                        hint = Fsp.hints.get(pn if pn is not None else n_src_dev)
                    # wgrad only feeds the optimiser: it runs on the side stream while dgrad / the previous
                    # block's BatchNorm backward continue on the main one
                    g_ready = torch.cuda.Event()
                    g_ready.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(g_ready)
                        # write-through target = the harness' flat gradient buffer, zeroed at the start of the step: the
                        # split reduction adds into it directly (no memset node per wgrad)
                        _lib.check(lib.wfsp_conv_wgrad_bf16(_lib.ptr(a16), n_in, _lib.ptr(n_src_dev), cin, _lib.ptr(g16),
                                                            n_dst, _lib.ptr(n_dst_dev), cout, _lib.ptr(pa), _lib.ptr(pb),
                                                            _lib.ptr(pn), kvol, pitch, hint, _lib.ptr(dw),
                                                            1 if (w_through and _write_through_zeroed) else 0, st()))
                    side_used = True
                    if not w_through:
                        grads[4 * bi] = dw.view(w_p.shape)
                if bi == 1 and grads_ready_hook is not None and _write_through > 0:
                    # blocks >= 1 are final once their wgrads (side stream) and BatchNorm gradients (main) have run
                    e_main, e_side = torch.cuda.Event(), torch.cuda.Event()
                    e_main.record(main)
                    e_side.record(side)
                    done = [q for q in params[4:] if q is not None and q.requires_grad]
                    # only if every one of them was written in place (a gradient handed back to autograd is accumulated
                    # after this function returns): no conv bias in these blocks, write-through targets attached
                    if (all(params[4 * j + 1] is None for j in range(1, len(blocks)))
                            and all(getattr(q, "_wfsp_grad_out", None) is not None for q in done)):
                        grads_ready_hook(done, [e_main, e_side])
                # ---- dgrad -> dy of the previous block (or of the stack's input)
                if bi > 0 or ctx.need_in_grad:
                    nbr_t = None if rb is None else (rb.nbr_out if conv.inverse else rb.nbr_in)
                    dy = torch.empty((n_in, cin), dtype=torch.float32, device=dev)
                    dy_parts = None
                    if n_in:
                        hint = Fsp.hints.get(n_src_dev)
                        # this dy arrives at the previous block's BatchNorm: its two reductions (sum dy', sum dy' xhat)
                        # are taken from the dgrad tile while it is on chip
                        bwd = None
                        if bi > 0 and blocks[bi - 1].bn is not None and _BWD_PARTS and cin <= 512 and not blocks[bi - 1].drop:
                            pb = blocks[bi - 1]
                            p_xf, p_mean, p_invstd = saved[bi - 1][1], saved[bi - 1][2], saved[bi - 1][3]
                            dy_parts = torch.empty((lib.wfsp_bn_partials_bytes(n_in, cin),), dtype=torch.uint8, device=dev)
                            bwd = (p_xf, p_mean, p_invstd, params[4 * (bi - 1) + 2], params[4 * (bi - 1) + 3], pb.relu, dy_parts)
                        ep = _lib.conv_epilogue(bwd=bwd)
                        _lib.check(lib.wfsp_conv_apply_bf16_ex(_lib.ptr(g16), n_dst, _lib.ptr(n_dst_dev), cout,
                                                               ctypes.c_void_p(wbuf.data_ptr() + offs[bi][1]), None,
                                                               _lib.ptr(nbr_t), kvol, _lib.ptr(dy), n_in, _lib.ptr(n_src_dev),
                                                               hint, cin, ctypes.byref(ep), st()))
            if side_used:
                joined = torch.cuda.Event()
                joined.record(side)
                main.wait_event(joined)
            del cross_stream  # main is ordered after every side-stream reader now
            d_feats = None
            if ctx.need_in_grad:
                d_feats = dy if ctx.in_dtype == torch.float32 else dy.to(ctx.in_dtype)
        return (None, None, None, d_feats) + tuple(grads)


def run(plan, x):
    """Runs the planned stack on SparseConvTensor x; returns a dense tensor (ToDense tail) or the
    output SparseConvTensor."""
    holder = {}
    out = FusedStackFunction.apply(plan, x, holder, x.features, *plan.params())
    if plan.to_dense:
        return out
    res = holder["tensor"]
    res.features = out
    return res
