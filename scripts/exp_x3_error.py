#!/usr/bin/env python
"""Actual error of the 3x3 sparse convolutions of the GEP stack (64 events) per math mode against float64: output,
input gradient, weight gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waveformml_b200 import spconv
from waveformml_b200.synth import make_events
dev = torch.device("cuda", 0)
ev = make_events(64, n_samples=1, seed=1234)
idx0 = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(dev)
torch.manual_seed(0)
pre = spconv.SparseConv2d(4, 4, 3, 1, 0, 1, 1, False).to(dev)
t1 = pre(spconv.SparseConvTensor(torch.rand(idx0.shape[0], 4, device=dev), idx0, [14, 11], 64))
for (cin, cout, idx, shape) in ((252, 158, idx0, [14, 11]), (158, 64, t1.indices, [12, 9])):
    layer = spconv.SparseConv2d(cin, cout, 3, 1, 0, 1, 1, False).to(dev)
    f = torch.rand(idx.shape[0], cin, device=dev)
    x = torch.zeros(64, cin, shape[0], shape[1], dtype=torch.float64, device=dev)
    x[idx[:, 0].long(), :, idx[:, 1].long(), idx[:, 2].long()] = f.double()
    x.requires_grad_(True)
    w = layer.weight.detach().double().requires_grad_(True)
    ref = torch.nn.functional.conv2d(x, w.permute(3, 2, 0, 1))
    for mode in ("fp32", "bf16", "bf16x3"):
        layer.math = mode
        layer.weight.grad = None
        fi = f.clone().requires_grad_(True)
        y = layer(spconv.SparseConvTensor(fi, idx, shape, 64))
        o = y.indices.long()
        r = ref[o[:, 0], :, o[:, 1], o[:, 2]]
        gen = torch.Generator(device="cpu").manual_seed(1)
        gy = torch.randn(tuple(y.features.shape), generator=gen).to(dev)
        y.features.backward(gy)
        if x.grad is not None:
            x.grad = None
        w.grad = None
        (r * gy.double()).sum().backward(retain_graph=True)
        dfr = x.grad[idx[:, 0].long(), :, idx[:, 1].long(), idx[:, 2].long()]
        e = lambda a, b: float((a.double() - b).norm() / b.norm())
        print("%d->%d rows %d->%d %-7s out %.3e  d_feats %.3e  d_weight %.3e" % (
            cin, cout, idx.shape[0], y.indices.shape[0], mode, e(y.features.detach(), r.detach()), e(fi.grad, dfr), e(layer.weight.grad, w.grad)))
