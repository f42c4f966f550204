"""Lightning-free replay of the reference's training step for the sparse-conv path, plus the
data-parallel plumbing (one process per GPU, events sharded by rank, one flat-gradient NCCL
all-reduce per step -- SURVEY.md 8e).

  LitPSD.training_step   src/engineering/LitPSD.py:94-104   predictions = model([c, f]); CE(mean)
  LitZ._process_batch    src/engineering/LitZ.py:89-107 + LitBase._calc_segment_loss
                         (src/engineering/LitBase.py:124-174): masked L1(sum) / N
  optimiser              config/examples/GEP.json:56-68 (SGD lr 0.02, momentum 0.98, nesterov)
  data parallel          src/utils/util.py:233-236 (Lightning DDPPlugin; plain BatchNorm1d, so
                         statistics stay rank-local)
"""
import torch
import torch.distributed as dist
from torch import nn

from . import spconv


def shard_events(n_events, rank, world):
    """Contiguous event range [lo, hi) owned by `rank`."""
    base, rem = divmod(n_events, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


def segment_l1_loss(indices, predictions, target, spatial_size, batch_size):
    """LitBase._calc_segment_loss with use_float=True, SE_only=False (LitBase.py:124-174): both the
    ones-mask and the target are densified through SparseConvTensor(...).dense()."""
    n = indices.shape[0]
    mask = spconv.SparseConvTensor(torch.ones((n, predictions.shape[1]), dtype=torch.float32, device=predictions.device),
                                   indices, spatial_size, batch_size).dense()
    tgt = target.unsqueeze(1) if target.dim() == 1 else target
    target_tensor = spconv.SparseConvTensor(tgt, indices, spatial_size, batch_size).dense()
    pred = mask * predictions
    return nn.functional.l1_loss(pred, target_tensor, reduction="sum") / n


class TrainStep:
    def __init__(self, model, task="psd", lr=0.02, momentum=0.98, nesterov=True, group=None):
        assert task in ("psd", "z")
        self.model, self.task, self.group = model, task, group
        self.grads = FlatGrads(model.parameters())
        self.opt = torch.optim.SGD(self.grads.params, lr=lr, momentum=momentum, nesterov=nesterov, foreach=True)
        self.criterion = nn.CrossEntropyLoss()

    def loss(self, indices, feats, target, batch_size):
        out = self.model([indices, feats, batch_size])
        if self.task == "psd":
            return self.criterion(out, target)
        return segment_l1_loss(indices, out, target, self.model.spatial_size, batch_size)

    def forward_backward(self, indices, feats, target, batch_size):
        self.grads.zero()
        loss = self.loss(indices, feats, target, batch_size)
        loss.backward()
        return loss

    def step(self, indices, feats, target, batch_size):
        loss = self.forward_backward(indices, feats, target, batch_size)
        self.grads.all_reduce_mean(self.group)
        self.opt.step()
        return loss.detach()
