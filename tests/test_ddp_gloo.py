"""Host-side logic of the data-parallel path on CPU with the gloo backend, world_size 2:
event sharding by rank and the single flat-gradient all-reduce.  The reference's
multi-GPU mode is plain DDP with rank-local BatchNorm (src/utils/util.py:233-236), so N ranks must
equal N micro-batches processed sequentially with their gradients averaged."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from waveformml_b200 import harness
from waveformml_b200.synth import make_events


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(12, 16), nn.BatchNorm1d(16), nn.ReLU(), nn.Linear(16, 3))


def _data(n_events):
    g = torch.Generator().manual_seed(1)
    return torch.randn(n_events, 12, generator=g), torch.randint(0, 3, (n_events,), generator=g)


def _worker(rank, world, port, n_events, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        grads = harness.FlatGrads(model.parameters())
        x, y = _data(n_events)
        lo, hi = harness.shard_events(n_events, rank, world)
        grads.zero()
        nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi]).backward()
        grads.all_reduce_mean()
        if rank == 0:
            torch.save(grads.flat.clone(), out)
    finally:
        dist.destroy_process_group()


def test_shard_events_partitions_the_batch():
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            spans = [harness.shard_events(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_synthetic_events_keep_event_ids_local():
    ev = make_events(10, n_samples=2, seed=3)
    lo, hi = harness.shard_events(10, 1, 2)
    rows = (ev["coords"][:, 2] >= lo) & (ev["coords"][:, 2] < hi)
    shard = ev["coords"][rows].copy()
    shard[:, 2] -= lo
    assert shard[:, 2].min() == 0 and shard[:, 2].max() == hi - lo - 1


def test_flat_gradient_allreduce_world2(tmp_path):
    n_events, world = 22, 2
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_events, out), nprocs=world, join=True)
    got = torch.load(out)
    # single process: the two shards one after the other, gradients averaged
    x, y = _data(n_events)
    ref = None
    for r in range(world):
        model = _model()
        grads = harness.FlatGrads(model.parameters())
        lo, hi = harness.shard_events(n_events, r, world)
        grads.zero()
        nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi]).backward()
        ref = grads.flat.clone() if ref is None else ref + grads.flat
    torch.testing.assert_close(got, ref / world, rtol=1e-6, atol=1e-7)


def test_flat_grads_are_views():
    model = _model()
    grads = harness.FlatGrads(model.parameters())
    assert grads.flat.numel() == sum(p.numel() for p in model.parameters())
    nn.functional.cross_entropy(model(torch.randn(5, 12)), torch.tensor([0, 1, 2, 0, 1])).backward()
    off = 0
    for p in model.parameters():
        assert p.grad.data_ptr() == grads.flat[off:off + p.numel()].data_ptr()  # autograd accumulated in place
        off += p.numel()
    assert float(grads.flat.abs().sum()) > 0
    grads.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in model.parameters())


def test_shard_events_by_rows_balances_hits():
    """Row-balanced contiguous sharding (load imbalance): a partition of the events, in order, whose
    per-rank row counts are within one event's rows of the ideal split."""
    import numpy as np
    rng = np.random.default_rng(5)
    for n, world in [(64, 2), (64, 8), (1024, 8), (10, 4), (3, 8), (8, 8), (1, 2)]:
        rows = np.clip(1 + rng.poisson(2.0, size=n), 1, 10)
        rows[: n // 4] *= 6  # a skewed head: equal event counts would overload rank 0
        spans = harness.shard_events_by_rows(rows, world)
        assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(lo <= hi for lo, hi in spans)
        if n >= world:
            assert all(hi > lo for lo, hi in spans)
            per = [int(rows[lo:hi].sum()) for lo, hi in spans]
            ideal = rows.sum() / world
            assert max(per) <= ideal + 2 * rows.max(), (n, world, per)
            by_events = [int(rows[lo:hi].sum()) for lo, hi in (harness.shard_events(n, r, world) for r in range(world))]
            assert max(per) <= max(by_events)
