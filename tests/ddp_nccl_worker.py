"""Worker of tests/test_gpu_ddp_nccl.py (launched with torch.distributed.run, one process per GPU).

Every rank trains captured steps of the PSD classifier on its own shard of a global batch (rank-local rulebooks and
BatchNorm; gradient exchange + SGD either as the fused peer-memory kernel wfsp_sgd_step_p2p or, with WFSP_P2P=0, as
NCCL all-reduce of gradient buckets + fused SGD, both inside the graph).  Rank 0 then
replays the same step WITHOUT communication -- shards fed sequentially through one process, gradients averaged --
which is the parity statement of SURVEY.md 8e, and checks (i) every rank ended with bit-identical parameters,
(ii) they equal the sequential-shard result."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from waveformml_b200 import batcher, harness, stacks
    from waveformml_b200.synth import make_events
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("WFSP_DDP_EVENTS", "48"))  # events per rank
    ev = make_events(B * world, n_samples=150, seed=4321)
    rows_of_event = torch.bincount(torch.from_numpy(ev["coords"][:, 2]).long(), minlength=B * world)
    starts = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(rows_of_event, 0)])

    def shard(r):
        lo, hi = harness.shard_events(B * world, r, world)
        r0, r1 = int(starts[lo]), int(starts[hi])
        c = torch.from_numpy(ev["coords"][r0:r1]).clone()
        c[:, 2] -= lo  # event ids are rank-local
        return c, torch.from_numpy(ev["wave"][r0:r1]), torch.from_numpy(ev["labels"][lo:hi])

    torch.manual_seed(0)
    model = stacks.PSDClassifier().to(dev).train()
    init = copy.deepcopy(model.state_dict())
    step = harness.GraphTrainStep(model, "psd", B, B * 10, 300, lr=0.02, momentum=0.98, nesterov=True)
    c, w, y = (t.to(dev) for t in shard(rank))
    step.load(c, w, y)
    try:
        step.capture()
        captured = True
    except Exception as exc:  # collective not capturable: keep the update outside the graph
        sys.stderr.write("rank %d: full-step capture failed (%s)\n" % (rank, exc))
        model.load_state_dict(init)
        step = harness.GraphTrainStep(model, "psd", B, B * 10, 300, capture_update=False)
        step.load(c, w, y)
        step.capture()
        captured = False
    loss = float(step.run())
    step.finish()  # the peer-memory exchange completes its closing barrier lazily (next replay or finish())
    flat = step.opt.flat_p.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok, msg = True, ""
    if rank == 0:
        for r in range(1, world):
            if not torch.equal(gathered[0], gathered[r]):
                ok, msg = False, "rank %d parameters differ from rank 0 (max |d| %.3e)" % (
                    r, float((gathered[0] - gathered[r]).abs().max()))
        # sequential-shard reference on this one process: same kernels, no communication
        ref = stacks.PSDClassifier().to(dev).train()
        ref.load_state_dict(init)
        rstep = harness.TrainStep(ref, "psd", lr=0.02, momentum=0.98, nesterov=True, data_parallel=False)
        acc = torch.zeros_like(rstep.grads.flat)
        for r in range(world):
            cr, wr, yr = (t.to(dev) for t in shard(r))
            idx, feats = batcher.pack_batch(cr, wr)
            bn_state = copy.deepcopy({k: v for k, v in ref.state_dict().items() if "running" in k or "num_batches" in k})
            rstep.forward_backward(idx, feats, yr, B)
            ref.load_state_dict(bn_state, strict=False)  # BatchNorm buffers are rank-local: irrelevant to the parameters
            acc += rstep.grads.flat
        rstep.grads.flat.copy_(acc)
        rstep.opt.step(1.0 / world)
        torch.cuda.synchronize()
        a, b = gathered[0].double(), rstep.opt.flat_p.detach().double()
        rel = float((a - b).norm() / b.norm())
        upd = float((b - torch.cat([init[k].reshape(-1).double() for k, _ in ref.named_parameters()])).norm() / b.norm())
        # same kernels on both sides; only the order of the cross-rank sum differs (fp32 all-reduce)
        if not (rel < 1e-6 and rel < 1e-3 * upd):
            ok, msg = False, "DDP parameters vs sequential shards: rel %.3e (update size %.3e)" % (rel, upd)
        print("ddp_nccl_parity world=%d events/rank=%d captured_allreduce=%s loss=%.6f rel_vs_sequential=%.3e "
              "update=%.3e ranks_identical=%s" % (world, B, captured, loss, rel, upd, ok or "differ" not in msg))
    # two more replays (momentum now non-zero, barrier epochs advance): the ranks must stay bit-identical
    for _ in range(2):
        step.run()
    step.finish()
    flat = step.opt.flat_p.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        for r in range(1, world):
            if not torch.equal(gathered[0], gathered[r]):
                ok, msg = False, "after 3 steps rank %d differs from rank 0" % r
        if not torch.isfinite(gathered[0]).all():
            ok, msg = False, "non-finite parameters after 3 steps"
        print("exchange=%s" % ("peer-memory kernel" if getattr(step.opt, "p2p", None) is not None else "NCCL buckets"))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0 and not ok:
        sys.stderr.write("FAIL: %s\n" % msg)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if int(flag.item()) == 1 else 1)  # no NCCL teardown under a captured graph (see bench.py)


if __name__ == "__main__":
    main()
