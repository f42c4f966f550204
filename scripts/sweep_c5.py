#!/usr/bin/env python
"""BASELINE.json configs[4]: high-occupancy (full-grid) events, sweep of batch size and channel width against
the reference CPU path.  For every (batch B, width w): the PSD classifier with the GEP channel ladder scaled to
w input channels (w/2 samples per PMT), one GPU; events/s of the graph-replayed training step (CUDA events, L2
flushed between steps) and of the CPU port of the reference algorithm (best thread count, a few steps; skipped
where one CPU step would take more than ~20 s).  Writes profiles/r1_sweep_C5.jsonl and .md.
Usage: sweep_c5.py [--cpu-max-rows N] [--no-cpu]   (--no-cpu: GPU side only; the CPU columns are carried over from
the committed profiles/r1_sweep_C5.jsonl, which the CPU port's speed does not depend on)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from waveformml_b200 import harness, stacks
from waveformml_b200.synth import make_events

dev = torch.device("cuda", 0)
cpu_max_rows = 40000
if "--cpu-max-rows" in sys.argv:
    cpu_max_rows = int(sys.argv[sys.argv.index("--cpu-max-rows") + 1])
no_cpu = "--no-cpu" in sys.argv
prev = {}
if no_cpu:
    for line in open(os.path.join(ROOT, "profiles", "r1_sweep_C5.jsonl")):
        r = json.loads(line)
        prev[(r["batch"], r["width"])] = r
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for w in (32, 64, 128, 256, 300):
    for B in (64, 256, 1024, 4096):
        torch.manual_seed(0)
        ch = stacks.gep_channel_ladder(w)
        model = stacks.PSDClassifier(n_samples=w // 2, channels=ch).to(dev).train()
        ev = make_events(B, n_samples=w // 2, seed=1234, full_grid=True)
        step = harness.GraphTrainStep(model, "psd", B, B * 154, w)
        c, wv, y = (torch.from_numpy(ev[k]).to(dev) for k in ("coords", "wave", "labels"))
        step.load(c, wv, y)
        step.capture()
        for _ in range(3):
            step.load(c, wv, y); step.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step.load(c, wv, y); step.run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        row = {"batch": B, "width": w, "channels": ch, "rows": int(c.shape[0]), "gpu_ms_per_step": ms,
               "gpu_events_per_s": B / (ms * 1e-3), "cpu_events_per_s": None, "cpu_threads": None}
        if no_cpu:
            p = prev.get((B, w), {})
            row["cpu_events_per_s"], row["cpu_threads"] = p.get("cpu_events_per_s"), p.get("cpu_threads")
        elif c.shape[0] <= cpu_max_rows:
            cstep = bench.cpu_step_fn(bench.cpu_reference_model(model), ev, model.n_linear)
            row["cpu_threads"] = bench.pick_cpu_threads(cstep)
            t0 = time.perf_counter(); n = 0
            while n < 2 or (time.perf_counter() - t0 < 3.0 and n < 20):
                cstep(); n += 1
            row["cpu_events_per_s"] = B * n / (time.perf_counter() - t0)
        out.append(row)
        print(json.dumps(row), flush=True)
        del step, model
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r1_sweep_C5.jsonl"), "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
with open(os.path.join(ROOT, "gpurun_out", "r1_sweep_C5.md"), "w") as f:
    f.write("| width | batch | rows | GPU ms/step | GPU events/s | CPU events/s (threads) | ratio |\n|---|---|---|---|---|---|---|\n")
    for r in out:
        cpu = "%.0f (%d)" % (r["cpu_events_per_s"], r["cpu_threads"]) if r["cpu_events_per_s"] else "-"
        ratio = "%.0fx" % (r["gpu_events_per_s"] / r["cpu_events_per_s"]) if r["cpu_events_per_s"] else "-"
        f.write("| %d | %d | %d | %.3f | %.0f | %s | %s |\n" % (r["width"], r["batch"], r["rows"], r["gpu_ms_per_step"],
                                                                r["gpu_events_per_s"], cpu, ratio))
