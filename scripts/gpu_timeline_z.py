#!/usr/bin/env python
"""Kernel timeline of one graph-replayed step of the z-regression model (BASELINE.json configs[2]).  Usage: [BATCH]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from waveformml_b200 import harness, stacks
from waveformml_b200.synth import make_events
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = stacks.ZRegressor().to(dev).train()
ev = make_events(B, n_samples=150, seed=4321)
c, w, z = (torch.from_numpy(ev[k]).to(dev) for k in ("coords", "wave", "z"))
step = harness.GraphTrainStep(model, "z", B, B * 10, 300)
step.load(c, w, z)
step.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_(); step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        flush.zero_(); step.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
fills = [i for i, e in enumerate(evs) if "FillFunctor<unsigned char>" in e.name]
seg = evs[fills[-1] + 1:]
t0 = seg[0].time_range.start
end = max(e.time_range.end for e in seg)
print("step span %.1f us, %d kernels, sum of durations %.1f us" % (end - t0, len(seg), sum(e.time_range.end - e.time_range.start for e in seg)))
for e in seg:
    print("%8.1f %7.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
