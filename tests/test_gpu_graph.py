"""The sync-free graph path (device-side row counts, capacity-sized buffers, our BatchNorm+ReLU
kernels, whole step captured in a CUDA graph) against the eager path and the CPU oracle."""
import copy

import pytest
import torch
from torch import nn

from oracle import mirror
from oracle import spconv_cpu as osp
from waveformml_b200 import batcher, harness, spconv, stacks
from waveformml_b200.spconv import functional as Fsp
from waveformml_b200.spconv import ops
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu


def _l2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-6 * b.numel() ** 0.5))


@pytest.mark.parametrize("n,c,relu", [(185, 252, True), (1000, 64, True), (37, 5, False), (1, 8, True), (300, 1, True)])
def test_batchnorm_relu_kernels_match_torch(cuda_device, n, c, relu):
    torch.manual_seed(n + c)
    cap = n + 77
    x = torch.randn(cap, c, device=cuda_device) * 3 + 1.5
    x[n:] = float("nan")  # rows past the live count must never be read
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    bn = nn.BatchNorm1d(c).to(cuda_device).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    ref_bn = copy.deepcopy(bn)
    xg = x.clone().requires_grad_(True)
    y = Fsp.batch_norm_relu(xg, n_dev, bn, relu)
    xr = x[:n].clone().requires_grad_(True)
    if n > 1:
        yr = ref_bn(xr)
    else:  # torch refuses a single row in training mode; the definition still holds (var = 0)
        yr = (xr - xr.mean(0)) * torch.rsqrt(xr.var(0, unbiased=False) + bn.eps) * ref_bn.weight + ref_bn.bias
    yr = torch.relu(yr) if relu else yr
    torch.testing.assert_close(y[:n], yr, rtol=1e-4, atol=1e-5)
    g = torch.randn(n, c, device=cuda_device)
    gfull = torch.full((cap, c), float("nan"), device=cuda_device)
    gfull[:n] = g
    y.backward(gfull)
    yr.backward(g)
    torch.testing.assert_close(xg.grad[:n], xr.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(bn.weight.grad, ref_bn.weight.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(bn.bias.grad, ref_bn.bias.grad, rtol=1e-3, atol=1e-4)
    if n > 1:
        torch.testing.assert_close(bn.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-6)
        assert int(bn.num_batches_tracked) == 1
    if n == 1:
        return
    # eval mode uses the running statistics
    bn.eval(), ref_bn.eval()
    ye = Fsp.batch_norm_relu(x, n_dev, bn, relu)
    yre = torch.relu(ref_bn(x[:n])) if relu else ref_bn(x[:n])
    torch.testing.assert_close(ye[:n], yre, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("k,s,p,subm", [(3, 1, 0, False), (3, 2, 1, False), (3, 1, 1, True), (5, 1, 2, True)])
def test_rulebook_with_device_count_matches_eager(cuda_device, k, s, p, subm):
    B = 23
    ev = make_events(B, n_samples=1, seed=4)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(cuda_device)
    n = idx.shape[0]
    cap = n + 100
    padded = torch.full((cap, 3), 12345, dtype=torch.int32, device=cuda_device)  # garbage past the live rows
    padded[:n] = idx
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    e = ops.build_rulebook(idx, B, [14, 11], [k, k], [s, s], [p, p], [1, 1], subm)
    g = ops.build_rulebook(padded, B, [14, 11], [k, k], [s, s], [p, p], [1, 1], subm, n_rows=n_dev)
    n_out = e.outids.shape[0]
    assert int(g.n_out_dev.item()) == n_out
    assert torch.equal(g.pair_num, e.pair_num)
    assert torch.equal(g.outids[:n_out], e.outids)
    assert torch.equal(g.pairs[:, :, :n], e.pairs)  # the capacity tail [n:] is unspecified (never read)
    assert torch.equal(g.nbr_out[:n_out], e.nbr_out) and torch.equal(g.nbr_in[:n], e.nbr_in)


@pytest.mark.parametrize("full_grid,B", [(False, 64), (True, 20), (True, 64)])
@pytest.mark.parametrize("subm", [False, True])
def test_rulebook_builder_choice_by_live_hint(cuda_device, full_grid, B, subm):
    """The expected-live-rows hint only picks the builder (one-CTA kernel up to 2,048 live rows, the phase
    kernels above); the rulebook is the same either way, also when the hint is wrong in both directions."""
    ev = make_events(B, n_samples=1, seed=9, full_grid=full_grid)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(cuda_device)
    n = idx.shape[0]
    cap = B * 154
    padded = torch.zeros((cap, 3), dtype=torch.int32, device=cuda_device)
    padded[:n] = idx
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    k, s, p = (3, 1, 1) if subm else (3, 2, 0)
    e = ops.build_rulebook(idx, B, [14, 11], [k, k], [s, s], [p, p], [1, 1], subm)
    n_out = e.outids.shape[0]
    for hint in (None, n, 1, cap):
        if hint is None:
            Fsp.hints.stop()
        else:
            Fsp.hints.start("replay")
            Fsp.hints.values = [hint]
        try:
            g = ops.build_rulebook(padded, B, [14, 11], [k, k], [s, s], [p, p], [1, 1], subm, n_rows=n_dev)
        finally:
            Fsp.hints.stop()
        assert int(g.n_out_dev.item()) == n_out
        assert torch.equal(g.pair_num, e.pair_num) and torch.equal(g.outids[:n_out], e.outids)
        assert torch.equal(g.pairs[:, :, :n], e.pairs)
        assert torch.equal(g.nbr_out[:n_out], e.nbr_out) and torch.equal(g.nbr_in[:n], e.nbr_in)


def _psd_inputs(B, seed, dev):
    ev = make_events(B, n_samples=150, seed=seed)
    return (torch.from_numpy(ev["coords"]).to(dev), torch.from_numpy(ev["wave"]).to(dev),
            torch.from_numpy(ev["labels"]).to(dev))


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16x3"])
def test_capacity_path_matches_eager_psd(cuda_device, mode):
    """Same weights, same batch: eager TrainStep (exact shapes, torch BatchNorm) vs the capacity-sized
    path with device-side counts (no graph yet) -- loss and every gradient."""
    spconv.set_math_mode(mode)
    try:
        B = 48
        torch.manual_seed(0)
        m1 = stacks.PSDClassifier().to(cuda_device).train()
        m2 = copy.deepcopy(m1)
        coords, wave, labels = _psd_inputs(B, 99, cuda_device)
        s1 = harness.TrainStep(m1, "psd")
        idx, feats = batcher.pack_batch(coords, wave)
        l1 = s1.forward_backward(idx, feats, labels, B)
        s2 = harness.GraphTrainStep(m2, "psd", B, B * 10, 300)
        s2.load(coords, wave, labels)
        idx2, feats2 = batcher.pack_batch(s2.coords, s2.wave, n_rows=s2.n_rows, tables=s2.tables)
        l2 = s2.forward_backward(idx2, feats2, s2.target, B, s2.n_rows)
        assert abs(float(l1.detach()) - float(l2.detach())) < 1e-4 * abs(float(l1.detach()))
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert _l2(b.grad, a.grad) < 2e-3, (k, _l2(b.grad, a.grad))
        for (k, a), (_, b) in zip(m1.named_buffers(), m2.named_buffers()):
            torch.testing.assert_close(b.float(), a.float(), rtol=1e-4, atol=1e-6, msg=k)
    finally:
        spconv.set_math_mode("bf16")


def test_graph_replay_tracks_eager_training(cuda_device):
    """One captured graph, several batches with different row counts: the loss trajectory and the
    final weights follow an eager run from the same initial state (fp32 math for a tight comparison)."""
    spconv.set_math_mode("fp32")
    try:
        B = 32
        torch.manual_seed(1)
        m1 = stacks.PSDClassifier().to(cuda_device).train()
        m2 = copy.deepcopy(m1)
        s1 = harness.TrainStep(m1, "psd", lr=0.01, momentum=0.9)
        s2 = harness.GraphTrainStep(m2, "psd", B, B * 10, 300, lr=0.01, momentum=0.9)
        batches = [_psd_inputs(B, 200 + i, cuda_device) for i in range(5)]
        s2.load(*batches[0])
        s2.capture()
        rows = set()
        for coords, wave, labels in batches:
            rows.add(coords.shape[0])
            idx, feats = batcher.pack_batch(coords, wave)
            le = float(s1.step(idx, feats, labels, B))
            s2.load(coords, wave, labels)
            lg = float(s2.run())
            assert s2.loss_value() == lg  # the pinned copy written by the step itself
            assert abs(le - lg) < 2e-3 * max(abs(le), 1e-3), (le, lg)
        assert len(rows) > 1  # the graph really was reused across different row counts
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert _l2(b, a) < 1e-3, (k, _l2(b, a))
    finally:
        spconv.set_math_mode("bf16")


def test_replay_after_foreign_gradient_writes(cuda_device):
    """The captured step has no fill launch (the optimiser's pass leaves the gradient buffer cleared); gradients
    written by anything else in between -- here an eager forward_backward on the same step object -- must not leak
    into the next replay."""
    spconv.set_math_mode("fp32")
    try:
        B = 16
        torch.manual_seed(3)
        m1 = stacks.PSDClassifier().to(cuda_device).train()
        m2 = copy.deepcopy(m1)
        s1 = harness.TrainStep(m1, "psd", lr=0.01, momentum=0.9)
        s2 = harness.GraphTrainStep(m2, "psd", B, B * 10, 300, lr=0.01, momentum=0.9)
        batches = [_psd_inputs(B, 300 + i, cuda_device) for i in range(3)]
        s2.load(*batches[0])
        s2.capture()
        for i, (coords, wave, labels) in enumerate(batches):
            idx, feats = batcher.pack_batch(coords, wave)
            s1.step(idx, feats, labels, B)
            if i == 1:  # gradients of some other batch land in the flat buffer between two replays
                bn_state = copy.deepcopy([b.clone() for b in m2.buffers()])
                oc, ow, ol = batches[2]
                oi, of = batcher.pack_batch(oc, ow)
                s2.forward_backward(oi, of, ol, B)
                assert float(s2.grads.flat.abs().sum()) > 0
                for b, saved in zip(m2.buffers(), bn_state):  # (the eager pass also moved the BatchNorm statistics)
                    b.copy_(saved)
            s2.load(coords, wave, labels)
            s2.run()
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert _l2(b, a) < 1e-3, (k, _l2(b, a))
        assert float(s2.grads.flat.abs().sum()) == 0.0  # left cleared by the optimiser's pass
    finally:
        spconv.set_math_mode("bf16")


def test_graph_path_vs_oracle_bf16(cuda_device):
    """The captured bf16 step against the CPU oracle with bf16-rounded GEMM operands."""
    B = 24
    torch.manual_seed(2)
    model = stacks.PSDClassifier().to(cuda_device).train()
    init = copy.deepcopy(model.state_dict())
    coords, wave, labels = _psd_inputs(B, 5, cuda_device)
    step = harness.GraphTrainStep(model, "psd", B, B * 10, 300, capture_update=False)
    step.load(coords, wave, labels)
    step.capture()
    step.graph.replay()  # forward + backward only (the optimiser is not in this graph): gradients inspectable
    loss = step.loss_out.clone()
    model_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    ref = stacks.PSDClassifier()
    ref.load_state_dict({k: v.cpu() for k, v in init.items()})
    osparse = mirror.to_oracle(ref.sparseModel).train()
    olinear = copy.deepcopy(ref.linear).train()
    osp.set_operand_rounding("bf16")
    try:
        idx, feats = batcher.pack_batch(coords, wave)
        d = mirror.run_stack(osparse, idx.cpu(), feats.cpu(), [14, 11], B)
        oloss = nn.CrossEntropyLoss()(olinear(d.view(-1, ref.n_linear)), labels.cpu())
        oloss.backward()
    finally:
        osp.set_operand_rounding(None)
    assert abs(float(loss) - float(oloss.detach())) < 1e-3 * abs(float(oloss.detach()))
    oparams = {"sparseModel." + k: v for k, v in osparse.named_parameters()}
    oparams.update({"linear." + k: v for k, v in olinear.named_parameters()})
    for k, g in model_grads.items():
        assert _l2(g, oparams[k].grad) < 1e-2, (k, _l2(g, oparams[k].grad))


def test_graph_path_z_regressor(cuda_device):
    """Masked-L1 segment loss (LitBase._calc_segment_loss) through the capacity path vs eager."""
    spconv.set_math_mode("fp32")
    try:
        B = 40
        torch.manual_seed(3)
        m1 = stacks.ZRegressor().to(cuda_device).train()
        m2 = copy.deepcopy(m1)
        ev = make_events(B, n_samples=150, seed=17)
        coords, wave = torch.from_numpy(ev["coords"]).to(cuda_device), torch.from_numpy(ev["wave"]).to(cuda_device)
        z = torch.from_numpy(ev["z"]).to(cuda_device)
        s1 = harness.TrainStep(m1, "z")
        idx, feats = batcher.pack_batch(coords, wave)
        l1 = s1.forward_backward(idx, feats, z, B)
        s2 = harness.GraphTrainStep(m2, "z", B, B * 10, 300, capture_update=False)
        s2.load(coords, wave, z)
        s2.capture()
        s2.graph.replay()
        l2 = s2.loss_out
        assert abs(float(l1.detach()) - float(l2)) < 1e-4 * abs(float(l1.detach()))
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert _l2(b.grad, a.grad) < 5e-3, (k, _l2(b.grad, a.grad))
    finally:
        spconv.set_math_mode("bf16")


def test_graph_path_reports_duplicate_inputs(cuda_device):
    """The captured step cannot raise on duplicate (event, x, y) rows; the flag of every captured rulebook is
    readable after a replay (dup_flag must reach the user)."""
    B = 8
    torch.manual_seed(5)
    model = stacks.PSDClassifier().to(cuda_device).train()
    coords, wave, labels = _psd_inputs(B, 11, cuda_device)
    step = harness.GraphTrainStep(model, "psd", B, B * 10, 300, capture_update=False)
    step.load(coords, wave, labels)
    step.capture()
    step.run()
    assert step.duplicate_inputs() is False
    bad = coords.clone()
    bad[1] = bad[0]  # two hits in the same cell of the same event
    step.load(bad, wave, labels)
    step.run()
    assert step.duplicate_inputs() is True
    step.load(coords, wave, labels)
    step.run()
    assert step.duplicate_inputs() is False


def test_two_backwards_accumulate_outside_write_through(cuda_device):
    """FlatGrads attaches write-through targets, but only TrainStep.forward_backward (one backward over zeroed
    gradients) may use them: two plain backwards must ACCUMULATE like torch.autograd."""
    B = 12
    torch.manual_seed(6)
    model = stacks.PSDClassifier().to(cuda_device).train()
    grads = harness.FlatGrads(model.parameters())
    coords, wave, labels = _psd_inputs(B, 13, cuda_device)
    idx, feats = batcher.pack_batch(coords, wave)
    crit = nn.CrossEntropyLoss()
    grads.zero()
    crit(model([idx, feats, B]), labels).backward()
    once = grads.flat.clone()
    crit(model([idx, feats, B]), labels).backward()
    twice = grads.flat.clone()
    # BatchNorm running statistics moved between the passes but training-mode outputs do not depend on them
    assert _l2(twice, 2 * once) < 1e-5
    # and the write-through step still produces the single-backward gradient
    step = harness.TrainStep(model, "psd")
    step.forward_backward(idx, feats, labels, B)
    assert _l2(step.grads.flat, once) < 1e-2  # bf16 math mode runs the head GEMMs of the harness step in TF32


@pytest.mark.parametrize("n,cap,C,Ct", [(3000, 3000, 1, 1), (777, 1200, 2, 1), (500, 500, 3, 3)])
def test_segment_l1_and_live_column_sums(cuda_device, n, cap, C, Ct):
    """The gather form of the masked L1 segment loss (wfsp_segment_l1_*) equals the reference's dense form
    (LitBase._calc_segment_loss) in value and gradient; wfsp_col_sum sums only the live rows of a capacity buffer."""
    from waveformml_b200.spconv import functional as Fsp
    dev = cuda_device
    B = 64
    g = torch.Generator().manual_seed(n)
    cells = torch.randperm(B * 14 * 11, generator=g)[:cap]
    idx = torch.stack([cells // 154, (cells % 154) // 11, cells % 11], 1).int().to(dev)
    n_dev = torch.tensor([n], dtype=torch.int32, device=dev) if cap != n else None
    pred = torch.randn(B, C, 14, 11, generator=g).to(dev).requires_grad_(True)
    tgt = torch.randn(cap, Ct, generator=g).to(dev)
    tgt_in = tgt[:, 0] if Ct == 1 else tgt
    loss = harness.segment_l1_loss(idx, pred, tgt_in, [14, 11], B, n_dev)
    (loss * 1.7).backward()
    # dense reference form on the live rows
    p2 = pred.detach().clone().requires_grad_(True)
    li = idx[:n].long()
    mask = torch.zeros(B, C, 14, 11, device=dev)
    mask[li[:, 0], :, li[:, 1], li[:, 2]] = 1.0
    td = torch.zeros(B, Ct, 14, 11, device=dev)
    td[li[:, 0], :, li[:, 1], li[:, 2]] = tgt[:n]
    ref = torch.nn.functional.l1_loss(mask * p2, td.expand(B, C, 14, 11), reduction="sum") / n
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    torch.testing.assert_close(pred.grad, p2.grad, rtol=1e-6, atol=1e-9)
    x = torch.randn(cap, 150, generator=g).to(dev)
    torch.testing.assert_close(Fsp.live_col_sum(x, n_dev), x[:n].sum(0), rtol=1e-4, atol=1e-4)
