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


class FlatGrads:
    """All parameter gradients live in one flat fp32 buffer (param.grad are views into it), so the
    data-parallel exchange is a single all-reduce of ~4 MB, issued once after backward."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=p0.device)
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
        the update instead of spending a kernel on the division."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0


class FlatSGD:
    """torch.optim.SGD(momentum, nesterov) semantics (config/examples/GEP.json:56-68) as ONE streaming kernel
    over flat buffers (wfsp_sgd_step).  The parameters are re-homed into one flat fp32 buffer (each
    `param.data` becomes a view of it, values preserved), laid out like the FlatGrads buffer."""

    def __init__(self, grads, lr, momentum=0.0, nesterov=False, weight_decay=0.0):
        self.grads, self.lr, self.momentum, self.nesterov, self.weight_decay = grads, lr, momentum, nesterov, weight_decay
        self.flat_p = torch.empty_like(grads.flat)
        off = 0
        with torch.no_grad():
            for p in grads.params:
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
        self.buf = torch.zeros_like(grads.flat)

    def step(self, grad_scale=1.0):
        from . import _lib
        lib = _lib.load()
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.wfsp_sgd_step(_lib.ptr(self.flat_p), _lib.ptr(self.grads.flat), _lib.ptr(self.buf),
                                         self.flat_p.numel(), float(self.lr), float(self.momentum), int(self.nesterov),
                                         float(self.weight_decay), float(grad_scale), _lib.stream()))

    def state_dict(self):
        return {"momentum_buffer": self.buf.clone()}

    def load_state_dict(self, sd):
        self.buf.copy_(sd["momentum_buffer"])


def segment_l1_loss(indices, predictions, target, spatial_size, batch_size, n_rows=None):
    """LitBase._calc_segment_loss with use_float=True, SE_only=False (LitBase.py:124-174): both the
    ones-mask and the target are densified through SparseConvTensor(...).dense()."""
    n = indices.shape[0]
    mask = spconv.SparseConvTensor(torch.ones((n, predictions.shape[1]), dtype=torch.float32, device=predictions.device),
                                   indices, spatial_size, batch_size, n_rows=n_rows).dense()
    tgt = target.unsqueeze(1) if target.dim() == 1 else target
    target_tensor = spconv.SparseConvTensor(tgt, indices, spatial_size, batch_size, n_rows=n_rows).dense()
    pred = mask * predictions
    denom = n if n_rows is None else n_rows.to(torch.float32)
    return nn.functional.l1_loss(pred, target_tensor, reduction="sum") / denom


class TrainStep:
    def __init__(self, model, task="psd", lr=0.02, momentum=0.98, nesterov=True, group=None, fused_head=True):
        assert task in ("psd", "z")
        self.model, self.task, self.group, self.fused_head = model, task, group, fused_head
        self.grads = FlatGrads(model.parameters())
        if self.grads.flat.is_cuda:
            self.opt = FlatSGD(self.grads, lr, momentum, nesterov)
        else:  # host-side tests of the plumbing (gloo): stock optimiser
            self.opt = torch.optim.SGD(self.grads.params, lr=lr, momentum=momentum, nesterov=nesterov, foreach=True)
        self.criterion = nn.CrossEntropyLoss()

    def _update(self):
        """gradient exchange + optimiser step"""
        if isinstance(self.opt, FlatSGD):
            self.opt.step(self.grads.all_reduce_sum(self.group))
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

    def forward_backward(self, indices, feats, target, batch_size, n_rows=None):
        self.grads.zero()
        # bf16 math mode = tensor-core operands with fp32 accumulation everywhere: the dense head's library
        # GEMMs (batches too large for the fused head) then run as TF32 tensor-core GEMMs instead of fp32 SIMT
        # ones (85 -> ~10 us for Linear(4480,116) at 1024 events).  fp32 mode keeps exact fp32 products.
        tf32 = self.grads.flat.is_cuda and spconv.get_math_mode() == "bf16"
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32 or prev
        try:
            # one backward over freshly zeroed gradients: the fused kernels write straight into the flat buffer
            with spconv.fused.grad_write_through():
                loss = self.loss(indices, feats, target, batch_size, n_rows)
                loss.backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return loss

    def step(self, indices, feats, target, batch_size, n_rows=None):
        loss = self.forward_backward(indices, feats, target, batch_size, n_rows)
        self._update()
        return loss.detach()


class GraphTrainStep(TrainStep):
    """One CUDA graph per model: static input buffers of `row_capacity` rows and `batch_size` events;
    `load()` copies a batch in (async from pinned host memory or device to device), `run()` replays
    the captured step.  Any batch with at most `row_capacity` rows and exactly `batch_size` events
    reuses the same graph -- the live row count is data, not shape."""

    def __init__(self, model, task, batch_size, row_capacity, n_chan, wave_dtype=torch.int16, scale=MAX_RANGE_INV,
                 capture_update=True, **kw):
        super().__init__(model, task, **kw)
        dev = self.grads.flat.device
        self.batch_size, self.row_capacity, self.scale = int(batch_size), int(row_capacity), scale
        self.coords = torch.zeros((row_capacity, 3), dtype=torch.int32, device=dev)
        self.wave = torch.zeros((row_capacity, n_chan), dtype=wave_dtype, device=dev)
        tshape = (batch_size,) if task == "psd" else (row_capacity,)
        self.target = torch.zeros(tshape, dtype=torch.int64 if task == "psd" else torch.float32, device=dev)
        self.n_rows = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._n_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.tables = batcher.item_tables([0, row_capacity], [0], dev)
        self.capture_update = capture_update
        self.graph = None
        self.loss_out = None
        self._stage, self._pending = None, None

    def load(self, coords, wave, target):
        """coords int32 [n,3] (x, y, event), wave [n,C], target [B] (psd) or [n] (z); host (pinned for an
        asynchronous copy) or device tensors."""
        n = coords.shape[0]
        if n > self.row_capacity:
            raise ValueError("batch has %d rows, graph capacity is %d" % (n, self.row_capacity))
        dev = self.coords.device
        if (coords.device == dev and wave.device == dev and target.device == dev and coords.dtype == self.coords.dtype
                and wave.dtype == self.wave.dtype and target.dtype == self.target.dtype and coords.is_contiguous()
                and wave.is_contiguous() and target.is_contiguous() and target.numel() <= self.target.numel()):
            # batch already in HBM: one launch stages all three buffers and the live row count
            self._stage_device(coords, wave, target, n)
            return
        self.coords[:n].copy_(coords, non_blocking=True)
        self.wave[:n].copy_(wave, non_blocking=True)
        self.target[:target.shape[0]].copy_(target, non_blocking=True)
        self.n_rows.fill_(n)  # the value travels as a kernel argument: no host buffer to race with

    def _stage_device(self, coords, wave, target, n):
        """One launch copies a device-resident batch into the graph's static buffers and sets the live row count."""
        lib = _lib.load()
        with torch.cuda.device(self.coords.device):
            _lib.check(lib.wfsp_stage_inputs(
                _lib.ptr(self.coords), _lib.ptr(coords), n * 3 * coords.element_size(),
                _lib.ptr(self.wave), _lib.ptr(wave), n * wave.shape[1] * wave.element_size(),
                _lib.ptr(self.target), _lib.ptr(target), target.numel() * target.element_size(),
                _lib.ptr(self.n_rows), n, _lib.stream()))

    def prefetch(self, coords, wave, target):
        """Double-buffered input staging (the DataLoader `pin_memory` + `non_blocking` pattern): the pinned host
        batch is copied to one of two device staging sets on a COPY stream, so the transfer of batch i+1 overlaps the
        compute of batch i; the next run() waits for the copy and moves it into the graph's buffers with one launch.
        At most one batch is pending; a staging set is reused only after the step that consumed it has read it."""
        n = coords.shape[0]
        if n > self.row_capacity:
            raise ValueError("batch has %d rows, graph capacity is %d" % (n, self.row_capacity))
        dev = self.coords.device
        if self._stage is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [{"coords": torch.empty_like(self.coords), "wave": torch.empty_like(self.wave),
                            "target": torch.empty_like(self.target), "ready": None, "free": None} for _ in range(2)]
            self._stage_i = 0
        slot = self._stage[self._stage_i]
        self._stage_i ^= 1
        cs = self._copy_stream
        if slot["free"] is not None:
            cs.wait_event(slot["free"])
        nt = target.shape[0]
        with torch.cuda.stream(cs):
            slot["coords"][:n].copy_(coords, non_blocking=True)
            slot["wave"][:n].copy_(wave, non_blocking=True)
            slot["target"][:nt].copy_(target, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(cs)
        slot["n"], slot["nt"] = n, nt
        self._pending = slot

    def _consume_prefetch(self):
        slot, self._pending = self._pending, None
        main = torch.cuda.current_stream()
        main.wait_event(slot["ready"])
        self._stage_device(slot["coords"], slot["wave"], slot["target"][:slot["nt"]], slot["n"])
        slot["free"] = torch.cuda.Event()
        slot["free"].record(main)

    def _body(self):
        # bf16 math + fused stack: the batcher writes the tensor-core operand format directly
        direct = spconv.get_math_mode() == "bf16" and spconv.fused.is_enabled()
        idx, feats = batcher.pack_batch(self.coords, self.wave, scale=self.scale, n_rows=self.n_rows,
                                        tables=self.tables, out_dtype=torch.bfloat16 if direct else torch.float32)
        loss = self.forward_backward(idx, feats, self.target, self.batch_size, self.n_rows)
        if self.capture_update:
            self._update()
        return loss.detach()

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
        self._restore(snap)

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        from . import _lib
        from .spconv.functional import hints
        with torch.cuda.stream(side):
            # warm-up on a side stream (allocator, cuBLAS handles, NCCL) before capture; the first pass also
            # records the live row counts of the loaded batch as launch-shape hints (see LaunchHints)
            for i in range(3):
                hints.start("record" if i == 0 else "replay")
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        lib = _lib.load()
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.wfsp_kernel_launches()
        hints.start("replay")
        from .spconv import ops as _ops
        del _ops.graph_dup_flags[:]
        try:
            with torch.cuda.graph(self.graph):
                self.loss_out = self._body()
        finally:
            hints.stop()
        self._dup_flags = list(_ops.graph_dup_flags)  # one int32 per captured rulebook, rewritten by every replay
        del _ops.graph_dup_flags[:]
        self.launches_per_replay = int(lib.wfsp_kernel_launches() - n0)  # libwfsp kernels in one replay

    def duplicate_inputs(self):
        """True if the batch of the LAST replay held duplicate (event, x, y) rows (a host readback: call it when
        validating data, not every step).  The eager path raises at rulebook construction instead."""
        flags = getattr(self, "_dup_flags", [])
        return bool(flags) and bool(torch.stack([f.reshape(()) for f in flags]).ne(0).any().item())

    def run(self):
        if self._pending is not None:
            self._consume_prefetch()
        if self.graph is None:
            self.capture()
        self.graph.replay()
        if not self.capture_update:
            self._update()
        return self.loss_out
