"""Fused dense head + cross-entropy for the PSD classifier (include/wfsp.h section 8).

    flatten(ToDense) -> Linear(k0, h1) -> Linear(h1, n_class) -> CrossEntropyLoss(mean)
    src/models/SPConvNet.py:67-68, src/models/ConvBlocks.py:82-102, src/engineering/LitPSD.py:94-104

as one autograd node of three launches (wfsp_head_ce_fwd = 2, wfsp_head_bwd = 1) instead of ~25 library
kernels.  Used by the training harness when the head is exactly two nn.Linear layers, the batch is small
(<= MAX_BATCH) and the loss is mean cross-entropy; anything else runs the stock torch modules.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib
from .spconv.fused import _grad_target

grads_ready_hook = None  # see spconv.fused.grads_ready_hook: called when the head's parameter gradients are final

# Set by a caller that promises to call join_deferred() on the backward stream before the gradients are used
# (harness.TrainStep.forward_backward): Linear-1's weight gradient then runs on a side stream, off the critical path.
defer_weight_grad = False
_deferred = []


def join_deferred():
    """Makes the current stream wait for every deferred weight-gradient product and releases its operands."""
    while _deferred:
        torch.cuda.current_stream().wait_event(_deferred.pop()[0])


MAX_BATCH = 256   # up to here the first Linear runs as our split-K launch + a one-CTA tail; beyond, Linear-1 is a library
                  # GEMM (TF32 in bf16 math mode) and the tail (Linear-2, loss, small backward half) one multi-CTA launch
_tickets = {}     # per device: the zero-initialised counter of wfsp_head_ce_tail (left zero by every call)
MAX_HIDDEN = 128
MAX_CLASSES = 64


def supported(linear, x, criterion=None):
    """linear: nn.Sequential of the head; x: [B, k0] input of the head."""
    if not (isinstance(linear, nn.Sequential) and len(linear) == 2 and all(isinstance(m, nn.Linear) for m in linear)):
        return False
    if criterion is not None and not (isinstance(criterion, nn.CrossEntropyLoss) and criterion.reduction == "mean"
                                      and criterion.weight is None and criterion.label_smoothing == 0.0
                                      and criterion.ignore_index < 0):
        return False
    l1, l2 = linear[0], linear[1]
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
            and l1.out_features <= MAX_HIDDEN and l2.out_features <= MAX_CLASSES and l1.in_features == x.shape[1]
            and l2.in_features == l1.out_features)


class HeadCEFunction(Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, labels):
        lib = _lib.load()
        x = x.contiguous()
        B, k0 = x.shape
        h1d, C = w1.shape[0], w2.shape[0]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        h1 = torch.empty((B, h1d), **f32)
        logits = torch.empty((B, C), **f32)
        loss = torch.empty((), **f32)
        dlogits = torch.empty((B, C), **f32)
        dh1 = torch.empty((B, h1d), **f32)
        dw2 = torch.empty((C, h1d), **f32)
        db2 = torch.empty((C,), **f32)
        labels = labels.contiguous()
        assert labels.dtype == torch.int64
        with torch.cuda.device(dev):
            if B <= MAX_BATCH:
                ws_bytes = lib.wfsp_head_workspace_bytes(B, k0, h1d)
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
                _lib.check(lib.wfsp_head_ce_fwd(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                                _lib.ptr(labels), B, k0, h1d, C, _lib.ptr(h1), _lib.ptr(logits),
                                                _lib.ptr(loss), _lib.ptr(dlogits), _lib.ptr(dh1), _lib.ptr(dw2),
                                                _lib.ptr(db2), _lib.ptr(ws), ws_bytes, _lib.stream()))
            else:
                if b1 is not None:
                    torch.addmm(b1, x, w1.t(), out=h1)
                else:
                    torch.mm(x, w1.t(), out=h1)
                key = (dev.type, dev.index)
                if key not in _tickets:
                    _tickets[key] = torch.zeros((1,), dtype=torch.int32, device=dev)
                ws_bytes = lib.wfsp_head_tail_workspace_bytes(B, h1d, C)
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
                _lib.check(lib.wfsp_head_ce_tail(_lib.ptr(h1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(labels), B, h1d, C,
                                                 _lib.ptr(logits), _lib.ptr(loss), _lib.ptr(dlogits), _lib.ptr(dh1),
                                                 _lib.ptr(dw2), _lib.ptr(db2), _lib.ptr(ws), ws_bytes,
                                                 _lib.ptr(_tickets[key]), _lib.stream()))
        ctx.save_for_backward(x, w1, dh1, dw2, db2)
        ctx.params = (w1, b1, w2, b2)
        ctx.logits = logits
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        x, w1, dh1, dw2_in, db2_in = ctx.saved_tensors
        w1_p, b1_p, w2_p, b2_p = ctx.params
        B, k0 = x.shape
        h1d, C = w1.shape[0], dw2_in.shape[0]
        dev = x.device
        go = grad_out.contiguous().float()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw1, w1_thr = _grad_target(w1_p, (h1d, k0), dev)
        db1, b1_thr = _grad_target(b1_p, (h1d,), dev) if b1_p is not None else (None, True)
        dw2, w2_thr = _grad_target(w2_p, (C, h1d), dev)
        db2, b2_thr = _grad_target(b2_p, (C,), dev) if b2_p is not None else (torch.empty((C,), device=dev), True)
        dh1s = torch.empty_like(dh1)
        with torch.cuda.device(dev):
            # small half in one launch; the two large products are plain GEMMs -> cuBLAS
            _lib.check(lib.wfsp_head_bwd_small(_lib.ptr(dh1), _lib.ptr(dw2_in), _lib.ptr(db2_in), _lib.ptr(go), B, h1d, C,
                                               _lib.ptr(dh1s), _lib.ptr(db1), _lib.ptr(dw2), _lib.ptr(db2), _lib.stream()))
            events = []
            if dx is not None:
                torch.mm(dh1s, w1, out=dx)
            if defer_weight_grad and dx is not None and w1_thr:  # (a gradient handed to autograd must be final on return)
                # the gradient of the stack below waits for dx only: Linear-1's weight gradient follows on a side stream,
                # beside the stack's backward pass (started behind dx: side by side the two small GEMMs slow each other)
                from .spconv.fused import _side_stream
                main, side = torch.cuda.current_stream(), _side_stream(dev, 5)
                fork = torch.cuda.Event()
                fork.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(fork)
                    torch.mm(dh1s.t(), x, out=dw1)
                    done = torch.cuda.Event()
                    done.record(side)
                # the operands stay referenced until join_deferred(): the allocator must not hand them out meanwhile
                _deferred.append((done, dh1s, x, dw1))
                events.append(done)
            else:
                torch.mm(dh1s.t(), x, out=dw1)
        if grads_ready_hook is not None and w1_thr and w2_thr and b1_thr and b2_thr:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            grads_ready_hook([q for q in (w1_p, b1_p, w2_p, b2_p) if q is not None], events + [ev])
        return (dx, None if w1_thr else dw1, None if b1_thr else db1, None if w2_thr else dw2, None if b2_thr else db2,
                None)


def head_cross_entropy(linear, x, labels):
    """mean cross-entropy of linear(x) against labels through the fused head."""
    l1, l2 = linear[0], linear[1]
    return HeadCEFunction.apply(x, l1.weight, l1.bias, l2.weight, l2.bias, labels)
