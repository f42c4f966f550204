#!/usr/bin/env python
"""Real (overlapped, warm-clock) kernel timeline of one graph-replayed training step through torch.profiler
(CUPTI): start offset, duration and stream of every kernel.  Usage: gpu_timeline.py [WORKLOAD] [BATCH]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from waveformml_b200 import harness, stacks
from waveformml_b200.synth import make_events

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
from waveformml_b200 import _lib
for kv in os.environ.get("WFSP_OPTIONS", "").split(","):  # e.g. WFSP_OPTIONS=apply_k_split=4,apply_split_stages=3
    if kv:
        k, v = kv.split("=")
        _lib.check(_lib.load().wfsp_set_option(k.encode(), int(v)))
torch.manual_seed(0)
model = stacks.PSDClassifier().to(dev).train()
batch = make_events(B, n_samples=150, seed=1234, full_grid=(wl == "C5"))
step = harness.GraphTrainStep(model, "psd", B, B * (154 if wl == "C5" else 10), 300)
c, w, y = (torch.from_numpy(batch[k]).to(dev) for k in ("coords", "wave", "labels"))
step.load(c, w, y)
step.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_(); step.load(c, w, y); step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        flush.zero_(); step.load(c, w, y); step.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last step = events after the last big fill
fills = [i for i, e in enumerate(evs) if "FillFunctor<unsigned char>" in e.name]
seg = evs[fills[-1] + 1:]
t0 = seg[0].time_range.start
end = max(e.time_range.end for e in seg)
print("step span %.1f us, %d kernels, sum of durations %.1f us" % (end - t0, len(seg), sum(e.time_range.end - e.time_range.start for e in seg)))
for e in seg:
    print("%8.1f %7.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
