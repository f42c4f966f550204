#!/usr/bin/env python
"""Kernel timeline of one graph-replayed DATA-PARALLEL step on rank 0 (launch with torch.distributed.run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from waveformml_b200 import harness, stacks
from waveformml_b200.synth import make_events

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 64
torch.manual_seed(0)
model = stacks.PSDClassifier().to(dev).train()
batch = make_events(B, n_samples=150, seed=1234 + rank)
step = harness.GraphTrainStep(model, "psd", B, B * 10, 300)
c, w, y = (torch.from_numpy(batch[k]).to(dev) for k in ("coords", "wave", "labels"))
step.load(c, w, y)
step.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.zero_(); step.run()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            flush.zero_(); step.run()
        torch.cuda.synchronize()
else:
    for _ in range(3):
        flush.zero_(); step.run()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    fills = [i for i, e in enumerate(evs) if "FillFunctor<unsigned char>" in e.name]
    seg = evs[fills[-1] + 1:]
    t0 = seg[0].time_range.start
    end = max(e.time_range.end for e in seg)
    print("world %d step span %.1f us, %d kernels" % (world, end - t0, len(seg)))
    for e in seg[-14:]:
        print("%8.1f %7.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:80]))
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
os._exit(0)
