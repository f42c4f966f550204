#!/usr/bin/env python
"""Where does bf16x3 deviate from fp32 inside the PSD classifier?  Activations per module and parameter gradients."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waveformml_b200 import spconv, stacks, harness, batcher
from waveformml_b200.synth import make_events
dev = torch.device("cuda", 0)
ev = make_events(64, n_samples=150, seed=1234)
coords, wave, labels = (torch.from_numpy(ev[k]).to(dev) for k in ("coords", "wave", "labels"))
idx, feats = batcher.pack_batch(coords, wave)
torch.manual_seed(0)
base = stacks.PSDClassifier().to(dev).train()
res = {}
for mode in ("fp32", "bf16x3"):
    spconv.set_math_mode(mode)
    m = copy.deepcopy(base)
    acts = []
    hooks = [mod.register_forward_hook(lambda mod, i, o, acts=acts: acts.append((type(mod).__name__, (o.features if hasattr(o, "features") else o).detach().double())))
             for mod in m.sparseModel]
    step = harness.TrainStep(m, "psd")
    loss = step.forward_backward(idx, feats, labels, 64)
    res[mode] = (acts, {k: p.grad.detach().double().clone() for k, p in m.named_parameters()}, float(loss))
spconv.set_math_mode("bf16")
a0, g0, l0 = res["fp32"]
a1, g1, l1 = res["bf16x3"]
print("loss", l0, l1)
for (n, x), (_, y) in zip(a0, a1):
    print("act %-14s rel %.3e" % (n, float((x - y).norm() / x.norm())))
for k in g0:
    print("grad %-24s rel %.3e" % (k, float((g0[k] - g1[k]).norm() / g0[k].norm().clamp_min(1e-30))))
