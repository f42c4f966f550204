"""Whole-model parity on the GPU: the reference's layer stacks forward + backward
through the training-step harness vs the same stacks run through the CPU oracle with shared
weights.  BatchNorm1d / ReLU / Linear are stock torch on both sides."""
import copy

import pytest
import torch
from torch import nn

from oracle import mirror
from oracle import spconv_cpu as osp
from waveformml_b200 import batcher, harness, spconv, stacks
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu


def _batch(B, ns, seed, dev, full=False):
    ev = make_events(B, n_samples=ns, seed=seed, full_grid=full)
    coords, wave = torch.from_numpy(ev["coords"]), torch.from_numpy(ev["wave"])
    idx, feats = batcher.pack_batch(coords.to(dev), wave.to(dev))
    return ev, idx, feats


def _rel(a, b):
    """Norm-wise relative error ||a - b||_2 / ||b||_2.  Whole-model gradients pass through ReLU gates
    and BatchNorm statistics computed from each side's own activations, so in bf16 mode a few gates
    flip and element-wise comparison is meaningless; the per-layer element-wise checks are in
    tests/test_gpu_conv.py."""
    a, b = a.detach().float().cpu().double(), b.detach().float().double()
    # a bias in front of BatchNorm has an analytically zero gradient: both sides hold only rounding
    # noise there, so the denominator is floored at 1e-4 per element
    return float((a - b).norm() / b.norm().clamp_min(1e-4 * b.numel() ** 0.5))


# (math mode, oracle operand rounding, gradient tolerance (norm-wise), loss tolerance (relative))
MODES = [("fp32", None, 1e-3, 1e-4),       # exact-fp32 kernels vs the fp32 oracle
         ("bf16x3", None, 2e-2, 1e-4),     # tcgen05 kernels with hi/lo-split operands vs the fp32 oracle.  Per layer they
                                           # are held to the fp32 tolerance (tests/test_gpu_conv.py; measured 4.5e-6
                                           # relative vs float64, fp32 kernels 4e-7, bf16 2.3e-3: scripts/exp_x3_error.py).
                                           # Whole-model GRADIENTS cross ReLU gates: a perturbation eps of the
                                           # activations flips a fraction ~eps of the gates, each flip is a full-size
                                           # error, so the norm-wise error goes like sqrt(eps) -- fp32 6e-4, bf16x3
                                           # 7e-3, bf16 8e-2 on this batch -- while loss and activations stay at eps
         ("bf16", "bf16", 1e-2, 1e-3),     # tcgen05 kernels vs the oracle with bf16-rounded GEMM operands
         ("bf16", None, 0.2, 1e-2)]        # tcgen05 kernels vs the fp32 oracle: a sanity cap only -- the bound that
                                           # is JUSTIFIED BY DATA is in tests/test_gpu_large.py
                                           # (test_bf16_vs_fp32_gap_is_the_operand_rounding: the oracle's own
                                           # gradients move by the same ~8 % when its operands are rounded to bf16)
MODE_IDS = ["fp32", "bf16x3-vs-fp32", "bf16-vs-emulated", "bf16-vs-fp32"]


@pytest.mark.parametrize("mode,rounding,tol,ltol", MODES, ids=MODE_IDS)
def test_psd_classifier_step_c1(cuda_device, mode, rounding, tol, ltol):
    """Config C1/C2: GEP stack, 64 events, CE loss; loss, logits and every parameter gradient."""
    torch.manual_seed(0)
    spconv.set_math_mode(mode)
    osp.set_operand_rounding(rounding)
    try:
        model = stacks.PSDClassifier().to(cuda_device).train()
        ev, idx, feats = _batch(64, 150, 1234, cuda_device)
        labels = torch.from_numpy(ev["labels"])
        step = harness.TrainStep(model, "psd")
        loss = step.forward_backward(idx, feats, labels.to(cuda_device), 64)
        # oracle twin
        osparse = mirror.to_oracle(model.sparseModel).train()
        olinear = copy.deepcopy(model.linear).cpu()
        for p in list(osparse.parameters()) + list(olinear.parameters()):
            p.grad = None
        d = mirror.run_stack(osparse, idx.cpu(), feats.cpu(), [14, 11], 64)
        ologits = olinear(d.view(-1, model.n_linear))
        oloss = nn.CrossEntropyLoss()(ologits, labels)
        oloss.backward()
        assert abs(float(loss.detach()) - float(oloss.detach())) / abs(float(oloss.detach())) < ltol
        gparams = dict(model.named_parameters())
        oparams = {"sparseModel." + k: v for k, v in osparse.named_parameters()}
        oparams.update({"linear." + k: v for k, v in olinear.named_parameters()})
        assert set(gparams) == set(oparams)
        for k in gparams:
            assert _rel(gparams[k].grad, oparams[k].grad) < tol, (k, _rel(gparams[k].grad, oparams[k].grad))
    finally:
        spconv.set_math_mode("bf16")
        osp.set_operand_rounding(None)


@pytest.mark.parametrize("mode,rounding,tol,ltol", MODES, ids=MODE_IDS)
def test_z_regressor_step(cuda_device, mode, rounding, tol, ltol):
    """Config C3 model (SingleEndedZCNN) with the masked-L1 segment loss of LitBase._calc_segment_loss."""
    torch.manual_seed(1)
    spconv.set_math_mode(mode)
    osp.set_operand_rounding(rounding)
    try:
        B = 96
        model = stacks.ZRegressor().to(cuda_device).train()
        ev, idx, feats = _batch(B, 150, 77, cuda_device)
        z = torch.from_numpy(ev["z"])
        step = harness.TrainStep(model, "z")
        loss = step.forward_backward(idx, feats, z.to(cuda_device), B)
        onet = mirror.to_oracle(model.model.network).train()
        pred = mirror.run_stack(onet, idx.cpu(), feats.cpu(), [14, 11], B)
        mask = osp.SparseConvTensor(torch.ones(idx.shape[0], 1), idx.cpu(), [14, 11], B).dense()
        tgt = osp.SparseConvTensor(z.unsqueeze(1), idx.cpu(), [14, 11], B).dense()
        oloss = nn.functional.l1_loss(mask * pred, tgt, reduction="sum") / idx.shape[0]
        oloss.backward()
        assert abs(float(loss.detach()) - float(oloss.detach())) / abs(float(oloss.detach())) < ltol
        for (k, a), (_, b) in zip(model.model.network.named_parameters(), onet.named_parameters()):
            assert a.shape == b.shape
            assert _rel(a.grad, b.grad) < tol, (k, _rel(a.grad, b.grad))
    finally:
        spconv.set_math_mode("bf16")
        osp.set_operand_rounding(None)


@pytest.mark.parametrize("name", ["ez_subm", "ioni_preserve"])
def test_other_stacks_forward(cuda_device, name):
    """SubM k5 with a shared rulebook key, and six conv/inverse-conv pairs (k3 and even k2)."""
    torch.manual_seed(2)
    spconv.set_math_mode("fp32")
    try:
        B = 40
        if name == "ez_subm":
            model, ns = stacks.EZSubM().to(cuda_device).eval(), 150
            net = model.network
        else:
            model, ns = stacks.IoniPreserve().to(cuda_device).eval(), 65
            net = model.model.func
        ev, idx, feats = _batch(B, ns, 5, cuda_device)
        with torch.no_grad():
            got = model([idx, feats, B])
            ref = mirror.run_stack(mirror.to_oracle(net).eval(), idx.cpu(), feats.cpu(), [14, 11], B)
        ref = ref.features if isinstance(ref, osp.SparseConvTensor) else ref
        assert got.shape == ref.shape
        assert _rel(got, ref) < 2e-3, _rel(got, ref)
    finally:
        spconv.set_math_mode("bf16")


def test_training_reduces_loss(cuda_device):
    """A few SGD steps on one batch reduce the loss (end-to-end sanity of fwd+bwd+update)."""
    torch.manual_seed(3)
    model = stacks.PSDClassifier().to(cuda_device).train()
    ev, idx, feats = _batch(64, 150, 1234, cuda_device)
    labels = torch.from_numpy(ev["labels"]).to(cuda_device)
    step = harness.TrainStep(model, "psd", lr=0.01, momentum=0.9)
    losses = [float(step.step(idx, feats, labels, 64)) for _ in range(8)]
    assert losses[-1] < losses[0], losses


def test_flat_sgd_matches_torch_nesterov(cuda_device):
    """wfsp_sgd_step over the flat buffers == torch.optim.SGD(momentum, nesterov) step for step
    (config/examples/GEP.json:56-68), including the 1 / world_size gradient scale."""
    from waveformml_b200 import harness
    torch.manual_seed(0)
    shapes = [(7, 5), (13,), (3, 3, 4, 6), (1,)]
    ours = [torch.nn.Parameter(torch.randn(s, device=cuda_device)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    fg = harness.FlatGrads(ours)
    opt = harness.FlatSGD(fg, lr=0.02, momentum=0.98, nesterov=True)
    ropt = torch.optim.SGD(ref, lr=0.02, momentum=0.98, nesterov=True)
    for step in range(4):
        gs = [torch.randn(s, device=cuda_device) for s in shapes]
        for p, r, g in zip(ours, ref, gs):
            p.grad.copy_(g)
            r.grad = (g * 0.5).clone()
        opt.step(grad_scale=0.5, zero_grads=step % 2 == 1)  # (265 elements: vector body + a scalar tail)
        ropt.step()
        for p, r in zip(ours, ref):
            torch.testing.assert_close(p.detach(), r.detach(), rtol=1e-5, atol=1e-6)
        if step % 2 == 1:
            assert float(fg.flat.abs().sum()) == 0.0  # optimizer.zero_grad() folded into the update pass
        else:
            assert float(fg.flat.abs().sum()) > 0.0
    assert all(p.data_ptr() >= opt.flat_p.data_ptr() for p in ours)  # parameters live in the flat buffer


@pytest.mark.parametrize("B,k0,h1,C", [(64, 4480, 116, 3), (5, 70, 128, 2), (200, 333, 17, 64), (1024, 4480, 116, 3),
                                       (300, 70, 33, 5), (257, 64, 128, 64)])
def test_fused_head_matches_torch(cuda_device, B, k0, h1, C):
    """wfsp_head_ce_fwd / wfsp_head_ce_tail / wfsp_head_bwd == Linear . Linear . CrossEntropyLoss(mean) of torch: loss,
    input gradient and all four parameter gradients, also under a non-unit incoming gradient.  Batches above 256 take
    the multi-CTA tail (Linear-1 as a library GEMM, exact fp32 here: TF32 is off outside the bf16 training step)."""
    from waveformml_b200 import head
    torch.manual_seed(B)
    lin = torch.nn.Sequential(torch.nn.Linear(k0, h1), torch.nn.Linear(h1, C)).to(cuda_device)
    x = torch.randn(B, k0, device=cuda_device, requires_grad=True)
    y = torch.randint(0, C, (B,), device=cuda_device)
    crit = torch.nn.CrossEntropyLoss()
    assert head.supported(lin, x, crit)
    loss = head.head_cross_entropy(lin, x, y)
    (loss * 0.7).backward()
    got = [loss.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in lin.parameters()]
    x.grad = None
    lin.zero_grad()
    ref = crit(lin(x), y)
    (ref * 0.7).backward()
    want = [ref.detach(), x.grad] + [p.grad for p in lin.parameters()]
    for a, b in zip(got, want):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-5 * max(float(b.abs().max()), 1e-3))


def test_head_bwd_all_in_one_entry(cuda_device):
    """wfsp_head_bwd (the C-ABI entry for callers without a GEMM library) against torch autograd."""
    from waveformml_b200 import _lib
    lib = _lib.load()
    B, k0, h1, C = 70, 150, 40, 5
    torch.manual_seed(3)
    x = torch.randn(B, k0, device=cuda_device)
    w1 = torch.randn(h1, k0, device=cuda_device, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    dh1 = torch.randn(B, h1, device=cuda_device)
    dw2_in, db2_in = torch.randn(C, h1, device=cuda_device), torch.randn(C, device=cuda_device)
    go = torch.tensor(1.3, device=cuda_device)
    dx, dw1, db1 = torch.empty_like(x), torch.empty_like(w1), torch.empty(h1, device=cuda_device)
    dw2, db2 = torch.empty_like(dw2_in), torch.empty_like(db2_in)
    with torch.cuda.device(cuda_device):
        _lib.check(lib.wfsp_head_bwd(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(dh1), _lib.ptr(dw2_in), _lib.ptr(db2_in), _lib.ptr(go),
                                     B, k0, h1, C, _lib.ptr(dx), _lib.ptr(dw1), _lib.ptr(db1), _lib.ptr(dw2), _lib.ptr(db2),
                                     _lib.stream()))
    ((xr @ w1.t()) * dh1).sum().mul(1.3).backward()
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dw1, w1.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(db1, dh1.sum(0) * 1.3, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dw2, dw2_in * 1.3)
    torch.testing.assert_close(db2, db2_in * 1.3)


@pytest.mark.parametrize("name", ["z", "ez_subm", "ioni_preserve"])
def test_c3_models_batch_1024_bf16(cuda_device, name):
    """BASELINE configs[2]: the regression / preserve stacks at batch 1024, training mode, tcgen05 bf16 kernels
    through the fused stack vs the oracle with bf16-rounded GEMM operands: output, and every parameter gradient
    of a sum-of-squares loss (norm-wise)."""
    torch.manual_seed(4)
    spconv.set_math_mode("bf16")
    osp.set_operand_rounding("bf16")
    try:
        B = 1024
        if name == "z":
            model, ns, net = stacks.ZRegressor().to(cuda_device).train(), 150, None
            net = model.model.network
        elif name == "ez_subm":
            model, ns = stacks.EZSubM().to(cuda_device).train(), 150
            net = model.network
        else:
            model, ns = stacks.IoniPreserve().to(cuda_device).train(), 65
            net = model.model.func
        ev, idx, feats = _batch(B, ns, 9, cuda_device)
        got = model([idx, feats, B])
        (got.square().sum() / got.numel()).backward()
        onet = mirror.to_oracle(net).train()
        ref = mirror.run_stack(onet, idx.cpu(), feats.cpu(), [14, 11], B)
        ref = ref.features if isinstance(ref, osp.SparseConvTensor) else ref
        (ref.square().sum() / ref.numel()).backward()
        assert got.shape == ref.shape and _rel(got, ref) < 1e-2, _rel(got, ref)
        # the preserve stack is 12 convolutions deep with 6 BatchNorm+ReLU gates between the loss and its first
        # layer: gates that flip on either side (each side thresholds its own bf16-rounded activations) compound,
        # so its early-layer gradients get a wider norm-wise bound than the 2-5 layer stacks
        gtol = 0.1 if name == "ioni_preserve" else 3e-2
        for (k, a), (_, b) in zip(net.named_parameters(), onet.named_parameters()):
            assert a.shape == b.shape
            assert _rel(a.grad, b.grad) < gtol, (k, _rel(a.grad, b.grad))
    finally:
        osp.set_operand_rounding(None)
