"""Lightning-free replay of the reference's training step for the sparse-conv path, plus the
data-parallel plumbing (one process per GPU, events sharded by rank, one flat-gradient NCCL
all-reduce per step -- SURVEY.md 8e).

  LitPSD.training_step   src/engineering/LitPSD.py:94-104   predictions = model([c, f]); CE(mean)
  LitZ._process_batch    src/engineering/LitZ.py:89-107 + LitBase._calc_segment_loss
                         (src/engineering/LitBase.py:124-174): masked L1(sum) / N
  optimiser              config/examples/GEP.json:56-68 (SGD lr 0.02, momentum 0.98, nesterov)
  data parallel          src/utils/util.py:233-236 (Lightning DDPPlugin; plain BatchNorm1d, so
                         statistics stay rank-local)

Two ways to run a step:
  TrainStep       eager: exact tensor shapes, stock torch BatchNorm1d / ReLU between our kernels, one
                  host readback per regular-conv rulebook (as upstream spconv needs for its shapes).
  GraphTrainStep  the whole step (batcher -> rulebooks -> fwd -> loss -> bwd -> all-reduce -> SGD) is
                  enqueued without any host readback -- row counts stay on the device, buffers are
                  capacity-sized -- captured once in a CUDA graph and replayed per batch.
"""
import os

import torch
import torch.distributed as dist
from torch import nn

from . import _lib, batcher, spconv
from .synth import MAX_RANGE_INV


def shard_events(n_events, rank, world):
    """Contiguous event range [lo, hi) owned by `rank`."""
    base, rem = divmod(n_events, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_events_by_rows(event_rows, world):
    """Contiguous event ranges balanced by ROWS (hits) instead of event counts -- SURVEY.md 8e: the conv work of a
    rank follows its rows, and hit multiplicities are skewed (1-10 hits per event in the reference's data).

    event_rows: rows of every event of the global batch, in batch order.  Returns `world` (lo, hi) event ranges
    that partition [0, n_events): boundary r is the event index where the running row count first reaches
    r / world of the total (every rank keeps at least one event while events remain).  Events stay whole and in
    order, so event ids stay rank-local after subtracting `lo`, exactly as with shard_events."""
    import numpy as np
    rows = np.asarray(event_rows, dtype=np.int64)
    n = int(rows.shape[0])
    cum = np.concatenate([[0], np.cumsum(rows)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        # the boundary that leaves the running count closest to the target
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, n)] - target):
            b -= 1
        b = max(b, bounds[-1] + (1 if bounds[-1] < n else 0))   # at least one event per rank while any remain
        b = min(b, n - (world - r)) if n >= world else min(b, n)  # ... and one left for every later rank
        bounds.append(max(b, bounds[-1]))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class PeerBuffer:
    """A tensor in peer-mapped (symmetric) memory: `tensor` is this rank's buffer, `ptrs_dev` a device array of the
    `world` ranks' pointers to theirs (torch.distributed._symmetric_memory does the IPC plumbing)."""

    def __init__(self, tensor, handle):
        self.tensor, self.handle = tensor, handle
        self.rank, self.world = handle.rank, handle.world_size
        self.ptrs_dev = torch.tensor(list(handle.buffer_ptrs), dtype=torch.int64, device=tensor.device)


def peer_buffer(numel, dtype, device, group=None):
    """PeerBuffer of `numel` elements, or None when not applicable (single process, CPU, WFSP_P2P=0, or symmetric
    memory unavailable -- the NCCL all-reduce path is used then)."""
    import os
    if group is False:
        return None
    if (device.type != "cuda" or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2
            or os.environ.get("WFSP_P2P", "1") == "0"):
        return None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        g = group if group is not None else dist.group.WORLD
        t = symm_mem.empty(max(int(numel), 4), dtype=dtype, device=device)
        h = symm_mem.rendezvous(t, g)
        return PeerBuffer(t[:numel], h)
    except Exception as exc:  # noqa: BLE001 -- any failure here just selects the NCCL path
        import sys
        sys.stderr.write("peer-mapped buffers unavailable (%s: %s); gradients go through NCCL all-reduce\n"
                         % (type(exc).__name__, exc))
        return None


class FlatGrads:
    """All parameter gradients live in one flat fp32 buffer (param.grad are views into it), so the
    data-parallel exchange is a single all-reduce of ~4 MB, issued once after backward."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        # data-parallel on GPUs: the buffer lives in peer-mapped memory, so the fused exchange + optimiser kernel
        # (wfsp_sgd_step_p2p) can read every rank's gradients over NVLink; otherwise an ordinary tensor
        self.peers = peer_buffer(total, torch.float32, p0.device, group)
        self.flat = self.peers.tensor if self.peers is not None else torch.zeros(total, dtype=torch.float32, device=p0.device)
        self.flat.zero_()
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            # inside spconv.fused.grad_write_through() (one backward per step over zeroed gradients) the fused
            # sparse stack and head write these gradients in place; outside it autograd accumulates as usual
            p._wfsp_grad_out = p.grad
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def all_reduce_sum(self, group=None):
        """SUM only; returns the factor (1 / world size) that turns it into the mean -- FlatSGD folds it into
        the update instead of spending a kernel on the division.  Ranges already exchanged during backward
        (early_reduce) are skipped; the side-stream exchanges are joined."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return 1.0
        done = sorted(getattr(self, "_early", []))
        pos, total = 0, self.flat.numel()
        for lo, hi, _ in done + [(total, total, None)]:
            if lo > pos:
                dist.all_reduce(self.flat[pos:lo], op=dist.ReduceOp.SUM, group=group)
            pos = max(pos, hi)
        main = torch.cuda.current_stream() if self.flat.is_cuda else None
        for _, _, ev in done:
            if ev is not None:
                main.wait_event(ev)
        self._early = []
        return 1.0 / dist.get_world_size(group)

    def offsets(self):
        if getattr(self, "_off", None) is None:
            self._off, off = {}, 0
            for p in self.params:
                self._off[id(p)] = (off, off + p.numel())
                off += p.numel()
        return self._off

    def early_reduce(self, params, events, group=None):
        """Called from backward (spconv.fused.grads_ready_hook / head.grads_ready_hook) when the gradients of `params`
        are final: if they form one contiguous range of the flat buffer, its all-reduce is issued NOW on a
        communication stream (ordered after `events`), overlapping the rest of the backward pass."""
        if not self.flat.is_cuda:
            return
        off = self.offsets()
        rng = sorted(off[id(p)] for p in params if id(p) in off)
        if not rng or any(rng[i][1] != rng[i + 1][0] for i in range(len(rng) - 1)):
            return  # not contiguous: left to the final exchange
        lo, hi = rng[0][0], rng[-1][1]
        if any(not (hi <= a or lo >= b) for a, b, _ in getattr(self, "_early", [])):
            return
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.flat.device)
        with torch.cuda.stream(self._comm):
            for ev in events:
                self._comm.wait_event(ev)
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._early = getattr(self, "_early", []) + [(lo, hi, done)]


class FlatSGD:
    """torch.optim.SGD(momentum, nesterov) semantics (config/examples/GEP.json:56-68) as ONE streaming kernel
    over flat buffers (wfsp_sgd_step).  The parameters are re-homed into one flat fp32 buffer (each
    `param.data` becomes a view of it, values preserved), laid out like the FlatGrads buffer."""

    def __init__(self, grads, lr, momentum=0.0, nesterov=False, weight_decay=0.0, group=None):
        self.grads, self.lr, self.momentum, self.nesterov, self.weight_decay = grads, lr, momentum, nesterov, weight_decay
        self.group = group
        dev = grads.flat.device
        # peer-mapped parameters + flags when the gradients are (all three or none: the ranks must agree)
        self.p2p = None
        if grads.peers is not None:
            pp = peer_buffer(grads.flat.numel(), torch.float32, dev, group)
            fl = peer_buffer(2 * grads.peers.world, torch.int32, dev, group)
            ok = torch.tensor([1 if (pp is not None and fl is not None) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                fl.tensor.zero_()
                self.p2p = {"params": pp, "flags": fl, "state": torch.zeros((4,), dtype=torch.int32, device=dev)}
                torch.cuda.synchronize(dev)
                dist.barrier(group=group)  # every rank's flags are zero before anyone starts a step
        self.flat_p = self.p2p["params"].tensor if self.p2p is not None else torch.empty_like(grads.flat)
        off = 0
        with torch.no_grad():
            for p in grads.params:
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
        self.buf = torch.zeros_like(grads.flat)

    def step_exchange(self, wait_now=True):
        """Peer-memory path: gradient reduce-scatter + update + parameter all-gather in one launch (every rank calls it
        once per step; no NCCL collective).  wait_now=False leaves the closing barrier to wait_exchange()."""
        from . import _lib
        lib = _lib.load()
        g, q = self.grads.peers, self.p2p
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.wfsp_sgd_step_p2p(
                _lib.ptr(self.flat_p), _lib.ptr(self.grads.flat), _lib.ptr(self.buf), self.flat_p.numel(), float(self.lr),
                float(self.momentum), int(self.nesterov), float(self.weight_decay), 1.0 / g.world, _lib.ptr(g.ptrs_dev),
                _lib.ptr(q["params"].ptrs_dev), _lib.ptr(q["flags"].ptrs_dev), _lib.ptr(q["flags"].tensor),
                _lib.ptr(q["state"]), g.rank, g.world, int(wait_now), _lib.stream()))
        self.exchange_pending = not wait_now

    def wait_exchange(self):
        """Completes the closing barrier of a step_exchange(wait_now=False): every rank's parameter stores have landed
        here and nobody reads this rank's gradients any more."""
        from . import _lib
        lib = _lib.load()
        q = self.p2p
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.wfsp_sgd_p2p_wait(_lib.ptr(q["flags"].tensor), _lib.ptr(q["state"]), self.grads.peers.world,
                                             _lib.stream()))
        self.exchange_pending = False

    def step(self, grad_scale=1.0, zero_grads=False):
        """zero_grads: the gradient buffer is cleared in the same pass (the next step needs no fill launch)."""
        from . import _lib
        lib = _lib.load()
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.wfsp_sgd_step_ex(_lib.ptr(self.flat_p), _lib.ptr(self.grads.flat), _lib.ptr(self.buf),
                                            self.flat_p.numel(), float(self.lr), float(self.momentum), int(self.nesterov),
                                            float(self.weight_decay), float(grad_scale), int(bool(zero_grads)),
                                            _lib.stream()))

    def state_dict(self):
        return {"momentum_buffer": self.buf.clone()}

    def load_state_dict(self, sd):
        self.buf.copy_(sd["momentum_buffer"])


class SegmentL1Function(torch.autograd.Function):
    """The masked L1 segment loss as one gather pass (wfsp_segment_l1_fwd / _bwd) instead of three ToDense scatters
    and ~15 element-wise kernels: inactive cells contribute |0 - 0| to the reference's dense l1_loss, so the loss is
    the row-wise L1 between the dense prediction at each hit's cell and the hit's target."""

    @staticmethod
    def forward(ctx, predictions, indices, target, n_rows, batch_size):
        lib = _lib.load()
        pred = predictions.contiguous()
        B, C, H, W = pred.shape
        n = indices.shape[0]
        tgt = (target.unsqueeze(1) if target.dim() == 1 else target).contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        ws = torch.empty((lib.wfsp_segment_l1_workspace_bytes(n),), dtype=torch.uint8, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(lib.wfsp_segment_l1_fwd(_lib.ptr(pred), _lib.ptr(indices), _lib.ptr(tgt), n, _lib.ptr(n_rows), C,
                                               tgt.shape[1], B, H, W, _lib.ptr(loss), _lib.ptr(ws), ws.numel(), _lib.stream()))
        ctx.save_for_backward(pred, indices, tgt)
        ctx.n_rows = n_rows
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        pred, indices, tgt = ctx.saved_tensors
        B, C, H, W = pred.shape
        d_pred = torch.zeros_like(pred)
        go = grad_out.contiguous().float()
        with torch.cuda.device(pred.device):
            _lib.check(lib.wfsp_segment_l1_bwd(_lib.ptr(pred), _lib.ptr(indices), _lib.ptr(tgt), indices.shape[0],
                                               _lib.ptr(ctx.n_rows), C, tgt.shape[1], B, H, W, _lib.ptr(go), _lib.ptr(d_pred),
                                               _lib.stream()))
        return d_pred, None, None, None, None


def segment_l1_loss(indices, predictions, target, spatial_size, batch_size, n_rows=None):
    """LitBase._calc_segment_loss with use_float=True, SE_only=False (LitBase.py:124-174): both the
    ones-mask and the target are densified through SparseConvTensor(...).dense().  On the GPU the same value is
    computed by one gather pass (SegmentL1Function); the dense formulation below is the CPU / reference form."""
    if (predictions.is_cuda and predictions.dim() == 4 and predictions.dtype == torch.float32 and indices.shape[1] == 3
            and indices.dtype == torch.int32 and indices.shape[0] > 0):
        tc = 1 if target.dim() == 1 else target.shape[1]
        if tc in (1, predictions.shape[1]):
            return SegmentL1Function.apply(predictions, indices.contiguous(), target, n_rows, batch_size)
    n = indices.shape[0]
    mask = spconv.SparseConvTensor(torch.ones((n, predictions.shape[1]), dtype=torch.float32, device=predictions.device),
                                   indices, spatial_size, batch_size, n_rows=n_rows).dense()
    tgt = target.unsqueeze(1) if target.dim() == 1 else target
    target_tensor = spconv.SparseConvTensor(tgt, indices, spatial_size, batch_size, n_rows=n_rows).dense()
    pred = mask * predictions
    denom = n if n_rows is None else n_rows.to(torch.float32)
    return nn.functional.l1_loss(pred, target_tensor, reduction="sum") / denom


# Linear-1's weight gradient of the fused head on a side stream (WFSP_TOPOLOGY=0: on the main stream, for A/B runs)
_HEAD_DEFER = os.environ.get("WFSP_TOPOLOGY", "1") != "0"
# the flat optimiser clears the gradient buffer in its own pass (no fill launch at the start of the next step)
_SGD_ZEROES = os.environ.get("WFSP_SGD_ZEROES", "1") != "0"


class TrainStep:
    def __init__(self, model, task="psd", lr=0.02, momentum=0.98, nesterov=True, group=None, fused_head=True,
                 data_parallel=True):
        """data_parallel=False: a purely local step even when torch.distributed is initialised (no peer-mapped buffers,
        which every rank would have to create together; _update() then must not be used across ranks)."""
        assert task in ("psd", "z")
        self.model, self.task, self.group, self.fused_head = model, task, group, fused_head
        self.grads = FlatGrads(model.parameters(), group if data_parallel else False)
        if self.grads.flat.is_cuda:
            self.opt = FlatSGD(self.grads, lr, momentum, nesterov, group=group if data_parallel else False)
        else:  # host-side tests of the plumbing (gloo): stock optimiser
            self.opt = torch.optim.SGD(self.grads.params, lr=lr, momentum=momentum, nesterov=nesterov, foreach=True)
        self.criterion = nn.CrossEntropyLoss()

    def _update(self, zero_grads=False):
        """gradient exchange + optimiser step (zero_grads: the flat optimiser also clears the gradient buffer)"""
        import os
        zero = bool(zero_grads)
        if os.environ.get("WFSP_NO_EXCHANGE") == "1" and isinstance(self.opt, FlatSGD):
            self.opt.step(1.0, zero)  # measurement aid: independent replicas, no gradient exchange at all
        elif isinstance(self.opt, FlatSGD) and self.opt.p2p is not None:
            # exchange + update over NVLink peer memory, no NCCL collective; the graph path completes the closing
            # barrier at the start of the NEXT replay (other ranks' stragglers hide behind its input handling)
            self.opt.step_exchange(wait_now=not getattr(self, "_defer_exchange_wait", False))
        elif isinstance(self.opt, FlatSGD):
            self.opt.step(self.grads.all_reduce_sum(self.group), zero)
        else:
            self.grads.all_reduce_mean(self.group)
            self.opt.step()

    def loss(self, indices, feats, target, batch_size, n_rows=None):
        x = [indices, feats, batch_size] if n_rows is None else [indices, feats, batch_size, n_rows]
        if self.task == "psd" and self.fused_head and hasattr(self.model, "forward_loss"):
            return self.model.forward_loss(x, target, self.criterion)
        out = self.model(x)
        if self.task == "psd":
            return self.criterion(out, target)
        return segment_l1_loss(indices, out, target, self.model.spatial_size, batch_size, n_rows)

    def forward_backward(self, indices, feats, target, batch_size, n_rows=None, zero=True, overlap_exchange=False):
        """loss + gradients.  overlap_exchange (only when _update() follows, i.e. from step()): gradient buckets that
        are final early are all-reduced during backward; a bare forward_backward never communicates."""
        if zero:
            self.grads.zero()
        # bf16 math mode = tensor-core operands with fp32 accumulation everywhere: the dense head's library
        # GEMMs (batches too large for the fused head) then run as TF32 tensor-core GEMMs instead of fp32 SIMT
        # ones (85 -> ~10 us for Linear(4480,116) at 1024 events).  fp32 mode keeps exact fp32 products.
        tf32 = self.grads.flat.is_cuda and spconv.get_math_mode() == "bf16"
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32 or prev
        from . import head as _head
        overlap = (overlap_exchange and self.grads.flat.is_cuda and isinstance(self.opt, FlatSGD)
                   and self.opt.p2p is None and dist.is_available()
                   and dist.is_initialized() and dist.get_world_size(self.group) > 1)
        if overlap:  # gradient buckets are exchanged as soon as they are final (head first, then the later conv blocks)
            hook = lambda params, events: self.grads.early_reduce(params, events, self.group)
            spconv.fused.grads_ready_hook, _head.grads_ready_hook = hook, hook
        try:
            # one backward over freshly zeroed gradients: the fused kernels write straight into the flat buffer
            with spconv.fused.grad_write_through(zeroed=True):
                loss = self.loss(indices, feats, target, batch_size, n_rows)
                self._after_forward(loss)
                # d(loss)/d(loss) = 1 from a persistent tensor: autograd would otherwise launch a fill for it every step
                if (getattr(self, "_one", None) is None or self._one.device != loss.device or self._one.dtype != loss.dtype
                        or self._one.shape != loss.shape):
                    self._one = torch.ones(loss.shape, dtype=loss.dtype, device=loss.device)
                _head.defer_weight_grad = self.grads.flat.is_cuda and _HEAD_DEFER
                loss.backward(gradient=self._one)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
            spconv.fused.grads_ready_hook, _head.grads_ready_hook = None, None
            _head.defer_weight_grad = False
            _head.join_deferred()
        self._grads_dirty = True
        return loss

    def _after_forward(self, loss):
        """Hook between the forward and the backward pass (GraphTrainStep copies the loss to the host from here)."""

    def step(self, indices, feats, target, batch_size, n_rows=None):
        loss = self.forward_backward(indices, feats, target, batch_size, n_rows, overlap_exchange=True)
        self._update()
        return loss.detach()


class GraphTrainStep(TrainStep):
    """One CUDA graph per input set: static input buffers of `row_capacity` rows and `batch_size` events;
    `load()` copies a batch in (async from pinned host memory or device to device), `run()` replays
    the captured step.  Any batch with at most `row_capacity` rows and exactly `batch_size` events
    reuses the same graph -- the live row count is data, not shape.

    n_buffers = 2 double-buffers the INPUTS: two input sets, each with its own captured graph over the same
    parameters / optimiser state, so `prefetch()` copies the next pinned host batch straight into the idle set on a
    copy stream while the current step runs -- no staging copy, no extra launch between the copy and the replay (the
    DataLoader `pin_memory` + `non_blocking` pattern of the reference's Lightning loop, src/utils/util.py:229-236)."""

    def __init__(self, model, task, batch_size, row_capacity, n_chan, wave_dtype=torch.int16, scale=MAX_RANGE_INV,
                 capture_update=True, n_buffers=1, **kw):
        super().__init__(model, task, **kw)
        dev = self.grads.flat.device
        self.batch_size, self.row_capacity, self.scale = int(batch_size), int(row_capacity), scale
        tshape = (batch_size,) if task == "psd" else (row_capacity,)
        self.sets = []
        for _ in range(max(1, int(n_buffers))):
            self.sets.append({
                "coords": torch.zeros((row_capacity, 3), dtype=torch.int32, device=dev),
                "wave": torch.zeros((row_capacity, n_chan), dtype=wave_dtype, device=dev),
                "target": torch.zeros(tshape, dtype=torch.int64 if task == "psd" else torch.float32, device=dev),
                "n_rows": torch.zeros((1,), dtype=torch.int32, device=dev),
                "n_host": torch.zeros((1,), dtype=torch.int32).pin_memory() if dev.type == "cuda" else None,
                "loss_host": torch.zeros((1,), dtype=torch.float32).pin_memory() if dev.type == "cuda" else None,
                "graph": None, "loss": None, "dup_flags": [], "ready": None, "done": None})
        self.cur = 0  # the set the next run() replays
        self.tables = batcher.item_tables([0, row_capacity], [0], dev)
        self.capture_update = capture_update
        self.loss_out = None
        self._cur_set, self._loss_copied = None, None
        self._stage, self._pending = None, None
        self._copy_stream = None

    # the first input set under the names the single-buffer interface always had
    coords = property(lambda self: self.sets[0]["coords"])
    wave = property(lambda self: self.sets[0]["wave"])
    target = property(lambda self: self.sets[0]["target"])
    n_rows = property(lambda self: self.sets[0]["n_rows"])
    graph = property(lambda self: self.sets[0]["graph"])

    def load(self, coords, wave, target, buf=None):
        """coords int32 [n,3] (x, y, event), wave [n,C], target [B] (psd) or [n] (z); host (pinned for an
        asynchronous copy) or device tensors.  Fills input set `buf` (default: the current one) and makes it the
        one the next run() replays."""
        n = coords.shape[0]
        if n > self.row_capacity:
            raise ValueError("batch has %d rows, graph capacity is %d" % (n, self.row_capacity))
        if buf is not None:
            self.cur = int(buf)
        st = self.sets[self.cur]
        dev = st["coords"].device
        if (coords.device == dev and wave.device == dev and target.device == dev and coords.dtype == st["coords"].dtype
                and wave.dtype == st["wave"].dtype and target.dtype == st["target"].dtype and coords.is_contiguous()
                and wave.is_contiguous() and target.is_contiguous() and target.numel() <= st["target"].numel()):
            # batch already in HBM: one launch stages all three buffers and the live row count
            self._stage_device(st, coords, wave, target, n)
            return
        st["coords"][:n].copy_(coords, non_blocking=True)
        st["wave"][:n].copy_(wave, non_blocking=True)
        st["target"][:target.shape[0]].copy_(target, non_blocking=True)
        st["n_rows"].fill_(n)  # the value travels as a kernel argument: no host buffer to race with

    def _stage_device(self, st, coords, wave, target, n):
        """One launch copies a device-resident batch into the set's static buffers and sets the live row count."""
        lib = _lib.load()
        with torch.cuda.device(st["coords"].device):
            _lib.check(lib.wfsp_stage_inputs(
                _lib.ptr(st["coords"]), _lib.ptr(coords), n * 3 * coords.element_size(),
                _lib.ptr(st["wave"]), _lib.ptr(wave), n * wave.shape[1] * wave.element_size(),
                _lib.ptr(st["target"]), _lib.ptr(target), target.numel() * target.element_size(),
                _lib.ptr(st["n_rows"]), n, _lib.stream()))

    def prefetch(self, coords, wave, target):
        """Queues the pinned host batch for the NEXT run() on a copy stream, so its transfer overlaps the step that
        is running.  With two input sets the copy lands directly in the idle set's buffers (after the last replay
        that read them has finished); with one set it goes through two staging sets and run() moves it into the
        graph's buffers with one launch.  At most one batch is pending."""
        n = coords.shape[0]
        if n > self.row_capacity:
            raise ValueError("batch has %d rows, graph capacity is %d" % (n, self.row_capacity))
        dev = self.sets[0]["coords"].device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        nt = target.shape[0]
        if len(self.sets) > 1:
            nxt = (self.cur + 1) % len(self.sets)
            st = self.sets[nxt]
            if st["done"] is not None:
                cs.wait_event(st["done"])  # the replay that last read this set
            st["n_host"][0] = n
            with torch.cuda.stream(cs):
                st["coords"][:n].copy_(coords, non_blocking=True)
                st["wave"][:n].copy_(wave, non_blocking=True)
                st["target"][:nt].copy_(target, non_blocking=True)
                st["n_rows"].copy_(st["n_host"], non_blocking=True)
                st["ready"] = torch.cuda.Event()
                st["ready"].record(cs)
            self._pending = ("set", nxt)
            return
        if self._stage is None:
            s0 = self.sets[0]
            self._stage = [{"coords": torch.empty_like(s0["coords"]), "wave": torch.empty_like(s0["wave"]),
                            "target": torch.empty_like(s0["target"]), "ready": None, "free": None} for _ in range(2)]
            self._stage_i = 0
        slot = self._stage[self._stage_i]
        self._stage_i ^= 1
        if slot["free"] is not None:
            cs.wait_event(slot["free"])
        with torch.cuda.stream(cs):
            slot["coords"][:n].copy_(coords, non_blocking=True)
            slot["wave"][:n].copy_(wave, non_blocking=True)
            slot["target"][:nt].copy_(target, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(cs)
        slot["n"], slot["nt"] = n, nt
        self._pending = ("stage", slot)

    def pending_ready(self):
        """Event after the H2D copies of the pending prefetch (the host may reuse its pinned buffers once it fired)."""
        kind, what = self._pending
        return self.sets[what]["ready"] if kind == "set" else what["ready"]

    def _consume_prefetch(self):
        (kind, what), self._pending = self._pending, None
        main = torch.cuda.current_stream()
        if kind == "set":
            self.cur = what
            main.wait_event(self.sets[what]["ready"])
            return
        slot = what
        main.wait_event(slot["ready"])
        self._stage_device(self.sets[self.cur], slot["coords"], slot["wave"], slot["target"][:slot["nt"]], slot["n"])
        slot["free"] = torch.cuda.Event()
        slot["free"].record(main)

    def finish(self):
        """Completes the gradient exchange of the last run() (peer-memory path defers its closing barrier to the next
        replay): call before reading parameters on the host / saving a checkpoint."""
        opt = self.opt
        if isinstance(opt, FlatSGD) and opt.p2p is not None:
            opt.wait_exchange()  # idempotent on the device
        torch.cuda.current_stream().synchronize()

    def _body(self, st=None):
        st = self.sets[self.cur] if st is None else st
        self._cur_set, self._loss_copied, self._zeroed = st, None, None
        p2p = isinstance(self.opt, FlatSGD) and self.opt.p2p is not None and self.capture_update
        self._defer_exchange_wait = p2p
        if p2p:
            # closing barrier of the PREVIOUS step's exchange (the kernel returns at once if none is pending)
            self.opt.wait_exchange()
        # bf16 math + fused stack: the batcher writes the tensor-core operand format directly
        direct = spconv.get_math_mode() == "bf16" and spconv.fused.is_enabled()
        # The rulebooks wait for the indices only, the first convolution for the features and the prepared weights:
        # the (larger) waveform conversion and the weight preparation run on side streams from the very start of the
        # step, the indices go first on the main stream (parallel branches of the captured graph).
        side = None
        if st["coords"].is_cuda:
            main = torch.cuda.current_stream()
            side = spconv.fused._side_stream(st["coords"].device, 2)
            # Layer 0's weight preparation (a 3 us launch on this stream) is the ONLY root of the captured step and
            # every side branch forks behind it: roots of parallel branches start staggered by up to 7 us in an order
            # that changes from replay to replay, and when the main chain came last the whole step was late.
            spconv.fused.prepare_stacks(self.model)
            side.wait_stream(main)
        # (measurement aid, profiles/r2_experiments.md: the waveform conversion on the main chain instead of a side
        # branch -- no gain, off)
        feats_main = os.environ.get("WFSP_FEATS_MAIN", "0") == "1"
        idx, feats = batcher.pack_batch(st["coords"], st["wave"], scale=self.scale, n_rows=st["n_rows"],
                                        tables=self.tables, out_dtype=torch.bfloat16 if direct else torch.float32,
                                        feats_stream=None if feats_main else side)
        if side is not None:
            if not feats_main:
                feats_ready = torch.cuda.Event()
                feats_ready.record(side)
                feats._wfsp_ready = feats_ready  # the first consumer on the main stream waits for it (fused.py)
            if getattr(self, "_grads_dirty", True):
                with torch.cuda.stream(side):
                    self.grads.zero()  # off the main stream too: the first gradient is written long after the join
                self._zeroed = torch.cuda.Event()
                self._zeroed.record(side)
        # With the flat optimiser in the step, ITS pass leaves the gradient buffer cleared for the next step: no fill
        # launch beside the first convolution (whose CTAs need whole SMs).  _grads_dirty tracks on the host whether
        # anything else (an eager forward_backward, a step without update) has written gradients since: the next
        # eager call / capture then zeroes explicitly, run() does so in front of a replay.
        self._sgd_zeroes = (self.capture_update and side is not None and isinstance(self.opt, FlatSGD)
                            and self.opt.p2p is None and _SGD_ZEROES)
        loss = self.forward_backward(idx, feats, st["target"], self.batch_size, st["n_rows"],
                                     zero=side is None and getattr(self, "_grads_dirty", True),
                                     overlap_exchange=self.capture_update)
        if self.capture_update:
            self._update(self._sgd_zeroes)
            if self._sgd_zeroes:
                self._grads_dirty = False
        loss = loss.detach()
        if self._loss_copied is not None:
            torch.cuda.current_stream().wait_event(self._loss_copied)  # join the copy branch (it finished long ago)
            self._loss_copied = None
        return loss

    def _after_forward(self, loss):
        # The loss lands in pinned host memory as part of the step (a copy node of the captured graph; the host reads it
        # after one stream synchronisation: loss_value()).  Issued on a side branch as soon as the forward pass has
        # produced it -- at the end of the step it was 7 us of the critical path behind the optimiser.
        st = self._cur_set
        if getattr(self, "_zeroed", None) is not None:
            # the gradient buffers were zeroed on a side branch (off the first convolution's dependencies): joined
            # here, in front of the first gradient write
            torch.cuda.current_stream().wait_event(self._zeroed)
            self._zeroed = None
        if st is None or st.get("loss_host") is None or not loss.is_cuda:
            return
        main = torch.cuda.current_stream()
        side = spconv.fused._side_stream(loss.device, 4)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            st["loss_host"].copy_(loss.detach().reshape(1), non_blocking=True)
            self._loss_copied = torch.cuda.Event()
            self._loss_copied.record(side)

    def loss_value(self):
        """Loss of the last run() as a Python float: waits for the step, reads the pinned copy the step wrote."""
        st = self.sets[self.cur]
        torch.cuda.current_stream().synchronize()
        if st.get("loss_host") is not None:
            return float(st["loss_host"][0])
        return float(self.loss_out.item())

    def _snapshot(self):
        import copy
        return (copy.deepcopy(self.model.state_dict()), copy.deepcopy(self.opt.state_dict()))

    def _restore(self, snap):
        # strictly in place: the captured graph holds the addresses of the parameters, BatchNorm buffers
        # and the momentum buffer
        self.model.load_state_dict(snap[0])
        if isinstance(self.opt, FlatSGD):
            self.opt.load_state_dict(snap[1])
            return
        old = snap[1]["state"]
        for i, p in enumerate(self.grads.params):
            st = self.opt.state.get(p, {})
            if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                prev = old.get(i, {}).get("momentum_buffer")
                if prev is None:
                    st["momentum_buffer"].zero_()  # == "no buffer yet" for SGD without dampening
                else:
                    st["momentum_buffer"].copy_(prev)

    def capture(self):
        snap = self._snapshot()  # the warm-up iterations below must not count as training steps
        self._capture()
        self.finish()  # peers may still be storing the last warm-up step's parameters into this rank's buffer
        if dist.is_available() and dist.is_initialized() and isinstance(self.opt, FlatSGD) and self.opt.p2p is not None:
            dist.barrier(group=self.group)
        self._restore(snap)
        if dist.is_available() and dist.is_initialized() and isinstance(self.opt, FlatSGD) and self.opt.p2p is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody starts a step before every rank has restored its parameters

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        from . import _lib
        from .spconv.functional import hints
        first = self.sets[self.cur]
        with torch.cuda.stream(side):
            # warm-up on a side stream (allocator, cuBLAS handles, NCCL) before capture; the first pass also
            # records the live row counts of the loaded batch as launch-shape hints (see LaunchHints)
            for i in range(3):
                hints.start("record" if i == 0 else "replay")
                self._body(first)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        lib = _lib.load()
        from .spconv import ops as _ops
        for st in self.sets:
            if st is not first:  # every set is captured over the batch the hints were recorded from
                for k in ("coords", "wave", "target", "n_rows"):
                    st[k].copy_(first[k])
            st["graph"] = torch.cuda.CUDAGraph()
            n0 = lib.wfsp_kernel_launches()
            hints.start("replay")
            del _ops.graph_dup_flags[:]
            try:
                # (measurement aid: capture the main chain on a high-priority stream -- no gain, off)
                prio = int(os.environ.get("WFSP_CAPTURE_PRIO", "0"))
                kw = {"stream": torch.cuda.Stream(priority=prio)} if prio else {}
                st["expects_clean"] = not getattr(self, "_grads_dirty", True)  # captured without a fill launch
                with torch.cuda.graph(st["graph"], **kw):
                    st["loss"] = self._body(st)
            finally:
                hints.stop()
            st["dup_flags"] = list(_ops.graph_dup_flags)  # one int32 per captured rulebook, rewritten by every replay
            del _ops.graph_dup_flags[:]
            self.launches_per_replay = int(lib.wfsp_kernel_launches() - n0)  # libwfsp kernels in one replay
        self.loss_out = first["loss"]

    def duplicate_inputs(self):
        """True if the batch of the LAST replay held duplicate (event, x, y) rows (a host readback: call it when
        validating data, not every step).  The eager path raises at rulebook construction instead."""
        flags = self.sets[self.cur]["dup_flags"]
        return bool(flags) and bool(torch.stack([f.reshape(()) for f in flags]).ne(0).any().item())

    def run(self):
        if self._pending is not None:
            self._consume_prefetch()
        st = self.sets[self.cur]
        if st["graph"] is None:
            self.capture()
        if st.get("expects_clean") and getattr(self, "_grads_dirty", False):
            # something outside the captured step wrote gradients since the optimiser last cleared them
            self.grads.zero()
            self._grads_dirty = False
        st["graph"].replay()
        if len(self.sets) > 1:
            st["done"] = torch.cuda.Event()
            st["done"].record(torch.cuda.current_stream())
        if not self.capture_update:
            self._update()
        self.loss_out = st["loss"]
        return self.loss_out
