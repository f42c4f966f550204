"""GPU sparse-conv layers (forward, dgrad, wgrad, bias grad) vs the CPU oracle on the same seeded
inputs and weights.

Tolerances (stated per north_star):
  fp32 mode  (CUDA-core FMA, exact fp32 products): rtol 1e-4, atol 1e-5 * max|ref|
  bf16 mode  (tcgen05, bf16 operands, fp32 accumulate): rtol 2e-2, atol 2e-2 * max|ref|
"""
import pytest
import torch

from oracle import mirror
from oracle import spconv_cpu as osp
from waveformml_b200 import spconv
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu

TOL = {"fp32": (1e-4, 1e-5), "bf16x3": (1e-4, 1e-5), "bf16": (2e-2, 2e-2), "bf16_emulated": (2e-3, 2e-3)}


def close(got, ref, mode, what=""):
    rtol, afrac = TOL[mode]
    got = got.detach().float().cpu()
    ref = ref.detach().float()
    atol = afrac * max(float(ref.abs().max()), 1e-30)
    torch.testing.assert_close(got, ref, rtol=rtol, atol=atol, msg=lambda m: "%s [%s]: %s" % (what, mode, m))


def l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def events(B, seed, C, full=False):
    ev = make_events(B, n_samples=1, seed=seed, full_grid=full)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(idx.shape[0], C, generator=g)  # waveform-like: non-negative
    return idx, feats


class Result:
    pass


def run_oracle(gnet, idx, feats, B, w, dense, rounding):
    osp.set_operand_rounding(rounding)
    try:
        onet = mirror.to_oracle(gnet)
        fo = feats.clone().requires_grad_(True)
        yo = onet(osp.SparseConvTensor(fo, idx, [14, 11], B))
        r = Result()
        r.indices = None if dense else yo.indices
        r.spatial = None if dense else list(yo.spatial_shape)
        yo = yo if dense else yo.features
        (yo * w).sum().backward()
        r.y, r.dfeats, r.params = yo, fo.grad, [p for p in onet.parameters()]
        return r
    finally:
        osp.set_operand_rounding(None)


def run_pair(layers, idx, feats, B, dev, mode, dense=False):
    """Runs the same stack on the GPU and through the oracle.  Returns the GPU result and a dict of
    oracle results: {"fp32": plain fp32 oracle, "bf16_emulated": oracle with bf16-rounded operands}
    (the latter only in bf16 mode)."""
    gnet = spconv.SparseSequential(*layers).to(dev)
    for m in gnet.modules():
        if isinstance(m, spconv.SparseConvolution):
            m.math = mode
    fg = feats.clone().to(dev).requires_grad_(True)
    yg = gnet(spconv.SparseConvTensor(fg, idx.to(dev), [14, 11], B))
    g = Result()
    g.indices = None if dense else yg.indices.cpu()
    g.spatial = None if dense else list(yg.spatial_shape)
    yg = yg if dense else yg.features
    gen = torch.Generator().manual_seed(7)
    w = torch.randn(tuple(yg.shape), generator=gen)
    (yg * w.to(dev)).sum().backward()
    g.y, g.dfeats, g.params = yg, fg.grad, [p for p in gnet.parameters()]
    refs = {"fp32": run_oracle(gnet, idx, feats, B, w, dense, None)}
    if mode == "bf16":
        refs["bf16_emulated"] = run_oracle(gnet, idx, feats, B, w, dense, "bf16")
    for r in refs.values():
        assert tuple(g.y.shape) == tuple(r.y.shape)
        if not dense:
            assert torch.equal(g.indices, r.indices) and g.spatial == r.spatial
    return g, refs


def check_all(g, refs, mode, name, elementwise_fp32_ref=True):
    """fp32 mode: element-wise vs the fp32 oracle.  bf16 mode: element-wise (tight) vs the oracle
    that rounds operands to bf16, and vs the plain fp32 oracle either element-wise with the stated
    bf16 tolerance (single layers) or norm-wise (stacks with ReLU gates in between)."""
    plan = [("fp32", mode)] if mode in ("fp32", "bf16x3") else [("bf16_emulated", "bf16_emulated")]
    if mode == "bf16" and elementwise_fp32_ref:
        plan.append(("fp32", "bf16"))
    for ref_name, tol in plan:
        r = refs[ref_name]
        close(g.y, r.y, tol, name + " out vs " + ref_name)
        close(g.dfeats, r.dfeats, tol, name + " d_features vs " + ref_name)
        for a, b in zip(g.params, r.params):
            close(a.grad, b.grad, tol, name + " d_param %s vs %s" % (tuple(a.shape), ref_name))
    if mode == "bf16" and not elementwise_fp32_ref:
        r = refs["fp32"]
        assert l2(g.y, r.y) < 2e-2 and l2(g.dfeats, r.dfeats) < 0.15
        for a, b in zip(g.params, r.params):
            assert l2(a.grad, b.grad) < 0.15, (tuple(a.shape), l2(a.grad, b.grad))


CASES = [
    # (name, layer factory, Cin)
    ("conv3_p0", lambda: [spconv.SparseConv2d(20, 24, 3, 1, 0, 1, 1, False)], 20),
    ("conv3_p1_bias", lambda: [spconv.SparseConv2d(12, 17, 3, 1, 1)], 12),
    ("conv3_s2", lambda: [spconv.SparseConv2d(10, 6, 3, 2, 1, 1, 1, True)], 10),
    ("conv2_even", lambda: [spconv.SparseConv2d(9, 5, 2, 1, 0, 1, 1, False)], 9),
    ("conv3_dil2", lambda: [spconv.SparseConv2d(8, 8, 3, 1, 2, 2)], 8),
    ("conv1x1", lambda: [spconv.SparseConv2d(30, 7, 1, 1, 0, 1, 1, True)], 30),
    ("conv1x1_cout1", lambda: [spconv.SparseConv2d(15, 1, 1, 1, 0)], 15),
    ("subm3", lambda: [spconv.SubMConv2d(16, 12, 3, 1, 1, indice_key="subm0")], 16),
    ("subm5_cout2", lambda: [spconv.SubMConv2d(11, 2, 5, 1, 2, indice_key="subm5")], 11),
    ("subm_chain_shared_key", lambda: [spconv.SubMConv2d(8, 8, 3, indice_key="subm0"),
                                       spconv.SubMConv2d(8, 4, 3, indice_key="subm0")], 8),
    ("conv_inverse", lambda: [spconv.SparseConv2d(6, 10, 3, 1, 1, 1, 1, False, indice_key="ind_0"),
                              spconv.SparseInverseConv2d(10, 10, 3, "ind_0", bias=False)], 6),
    ("conv_inverse_k2", lambda: [spconv.SparseConv2d(7, 5, 2, 1, 0, 1, 1, False, indice_key="ind_2"),
                                 spconv.SparseInverseConv2d(5, 5, 2, "ind_2", bias=True)], 7),
    ("wide_odd_channels", lambda: [spconv.SparseConv2d(158, 64, 3, 1, 0, 1, 1, False)], 158),
    ("cout_gt_256", lambda: [spconv.SparseConv2d(36, 300, 3, 1, 1, 1, 1, True)], 36),
    ("cin_300_k1", lambda: [spconv.SparseConv2d(300, 252, 1, 1, 0, 1, 1, False)], 300),
]


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16_per_layer", "bf16x3"])
@pytest.mark.parametrize("name,factory,cin", CASES, ids=[c[0] for c in CASES])
def test_layer_parity(cuda_device, name, factory, cin, mode):
    """bf16: the stack runs as one fused node with bf16-resident operands (spconv/fused.py);
    bf16_per_layer: the same kernels module by module through the fp32-in / fp32-out C entries;
    bf16x3: the tensor-core kernels with hi/lo-split operands, held to the fp32 tolerance."""
    torch.manual_seed(sum(name.encode()))
    B = 19
    idx, feats = events(B, 21, cin)
    spconv.set_fused(mode != "bf16_per_layer")
    try:
        mode = "bf16" if mode == "bf16_per_layer" else mode
        g, refs = run_pair(factory(), idx, feats, B, cuda_device, mode)
    finally:
        spconv.set_fused(True)
    check_all(g, refs, mode, name)


@pytest.mark.parametrize("B,full,bn_in_conv", [(40, False, 0), (160, True, 0), (40, False, 1), (100, False, 1)],
                         ids=["small", "24640_rows", "small_bn_in_conv", "mid_bn_in_conv"])
def test_fused_equals_per_layer(cuda_device, B, full, bn_in_conv):
    """The fused stack rounds the same values to bf16 at the same points as the per-layer path, so
    a conv-BN-ReLU stack must agree with it closely (outputs, input and parameter gradients, BN buffers).
    The large case crosses the row count above which BatchNorm statistics come from the convolution
    epilogue (per-32-row partials) and the multi-kernel BatchNorm / rulebook paths are taken."""
    idx, feats = events(B, 23, 30, full=full)

    def make():
        torch.manual_seed(11)
        return spconv.SparseSequential(
            spconv.SparseConv2d(30, 26, 1, 1, 0, 1, 1, False), torch.nn.BatchNorm1d(26), torch.nn.ReLU(),
            spconv.SparseConv2d(26, 20, 3, 1, 1, 1, 1, True, indice_key="k"), torch.nn.ReLU(),
            spconv.SparseInverseConv2d(20, 12, 3, "k", bias=False),
            # bias=False before BatchNorm: its gradient is identically zero in exact arithmetic (pure round-off)
            spconv.SubMConv2d(12, 9, 3, bias=False, indice_key="subm0"), torch.nn.BatchNorm1d(9),
            spconv.ToDense()).to(cuda_device)

    from waveformml_b200 import _lib
    res = []
    for fused_on in (True, False):
        spconv.set_fused(fused_on)
        # bn_in_conv: the BatchNorm behind a small convolution launch is finished inside that launch (grid barrier;
        # option apply_bn_fuse, off by default) -- same results as the stand-alone BatchNorm launch
        _lib.check(_lib.load().wfsp_set_option(b"apply_bn_fuse", bn_in_conv if fused_on else 0))
        try:
            net = make()
            f = feats.clone().to(cuda_device).requires_grad_(True)
            y = net(spconv.SparseConvTensor(f, idx.to(cuda_device), [14, 11], B))
            gen = torch.Generator().manual_seed(5)
            w = torch.randn(tuple(y.shape), generator=gen).to(cuda_device)
            (y * w).sum().backward()
            res.append((y.detach(), f.grad, [p.grad for p in net.parameters()],
                        [b.clone() for b in net.buffers()]))
        finally:
            spconv.set_fused(True)
            _lib.check(_lib.load().wfsp_set_option(b"apply_bn_fuse", 0))
    (ya, fa, pa, ba), (yb, fb, pb, bb) = res
    # not bit-equal: the eager per-layer path normalises with torch's BatchNorm1d, whose statistics differ
    # from ours in the last fp32 bits, which can flip an occasional bf16 rounding of an activation
    def close_(a, b):
        torch.testing.assert_close(a, b, rtol=5e-3, atol=5e-3 * max(float(b.abs().max()), 1e-6))
    close_(ya, yb)
    close_(fa, fb)
    for a, b in zip(pa, pb):
        close_(a, b)
    for a, b in zip(ba, bb):  # BatchNorm running statistics and step counters
        torch.testing.assert_close(a.float(), b.float(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_gep_conv_stack_dense(cuda_device, mode):
    """The three convolutions of the GEP stack + ReLU + ToDense (no BN so the comparison isolates our
    kernels).  In bf16 mode the gradients cross ReLU gates taken from each side's own activations, so
    against the fp32 oracle they are compared norm-wise; against the bf16-emulating oracle element-wise."""
    torch.manual_seed(3)
    B = 64
    ev = make_events(B, n_samples=150, seed=1234)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()
    feats = torch.from_numpy(ev["wave"]).float() / 16383.0
    layers = [spconv.SparseConv2d(300, 252, 1, 1, 0, 1, 1, False), torch.nn.ReLU(),
              spconv.SparseConv2d(252, 158, 3, 1, 0, 1, 1, False), torch.nn.ReLU(),
              spconv.SparseConv2d(158, 64, 3, 1, 0, 1, 1, False), spconv.ToDense()]
    g, refs = run_pair(layers, idx, feats, B, cuda_device, mode, dense=True)
    assert tuple(g.y.shape) == (64, 64, 10, 7)
    check_all(g, refs, mode, "gep stack", elementwise_fp32_ref=False)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_multi_tile_rows(cuda_device, mode):
    """More than one 128-row tile, many CTAs, split wgrad reduction."""
    torch.manual_seed(5)
    B = 300
    idx, feats = events(B, 31, 40)
    g, refs = run_pair([spconv.SparseConv2d(40, 48, 3, 1, 1, 1, 1, True)], idx, feats, B, cuda_device, mode)
    check_all(g, refs, mode, "multi tile")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_big_kernel_unstaged_neighbours(cuda_device, mode):
    """Kernel volume 49 > 32: the CTA's neighbour tile is not staged in shared memory."""
    torch.manual_seed(6)
    B = 9
    idx, feats = events(B, 33, 12)
    g, refs = run_pair([spconv.SubMConv2d(12, 10, 7, indice_key="subm7")], idx, feats, B, cuda_device, mode)
    check_all(g, refs, mode, "subm7")


def test_linearity_full_size(cuda_device):
    """Size-independent property at BASELINE full size (C5, 1024 full-grid events): the layer is
    linear in its input, conv(a + 2b) == conv(a) + 2 conv(b), and 1x1 == dense matmul."""
    B = 1024
    ev = make_events(B, n_samples=1, seed=2, full_grid=True)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(cuda_device)
    N = idx.shape[0]
    torch.manual_seed(0)
    a = torch.rand(N, 64, device=cuda_device)
    b = torch.rand(N, 64, device=cuda_device)
    conv = spconv.SparseConv2d(64, 32, 3, 1, 0, 1, 1, False).to(cuda_device)
    for mode, tol in (("fp32", 1e-4), ("bf16", 3e-2)):
        conv.math = mode
        with torch.no_grad():
            ya = conv(spconv.SparseConvTensor(a, idx, [14, 11], B)).features
            yb = conv(spconv.SparseConvTensor(b, idx, [14, 11], B)).features
            yab = conv(spconv.SparseConvTensor(a + 2 * b, idx, [14, 11], B)).features
        assert ya.shape == (1024 * 108, 32)
        err = (yab - (ya + 2 * yb)).abs().max() / yab.abs().max()
        assert float(err) < tol, (mode, float(err))
    pw = spconv.SparseConv2d(64, 48, 1, 1, 0, 1, 1, False).to(cuda_device)
    pw.math = "fp32"
    with torch.no_grad():
        y = pw(spconv.SparseConvTensor(a, idx, [14, 11], B)).features
        ref = a @ pw.weight.view(64, 48)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)


def test_half_precision_features(cuda_device):
    """system_config.half_precision feeds fp16 features (src/datasets/HDF5Dataset.py:227): accepted,
    computed in fp32/bf16 internally, returned in the input dtype."""
    idx, feats = events(5, 3, 8)
    conv = spconv.SubMConv2d(8, 4, 3, indice_key="subm0").to(cuda_device)
    y = conv(spconv.SparseConvTensor(feats.half().to(cuda_device), idx.to(cuda_device), [14, 11], 5))
    assert y.features.dtype == torch.float16 and y.features.shape == (idx.shape[0], 4)


def test_wgrad_launch_hint_never_drops_pairs(cuda_device):
    """Graph path: the launch-shape hint (pairs of the fullest offset seen while warming up) may be far
    below the live pair count of a later batch; the split of the pair list must still cover every pair."""
    from waveformml_b200.spconv import functional as Fsp
    from waveformml_b200.spconv import ops
    B = 300
    idx, feats = events(B, 41, 24)
    idx_d = idx.to(cuda_device)
    n = idx.shape[0]
    out_idx, pairs, pair_num = ops.get_indice_pairs(idx_d, B, [14, 11], [3, 3], [1, 1], [1, 1], [1, 1], subm=True)
    g = torch.Generator().manual_seed(1)
    dout = torch.randn(n, 16, generator=g).to(cuda_device)
    fd = feats.to(cuda_device)
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    ref = Fsp.conv_wgrad(fd, dout, pairs[0], pairs[1], pair_num, 9, "bf16")
    Fsp.hints.start("replay")
    Fsp.hints.values = [64]  # "expect 64 pairs per offset": the real count is ~10x that
    try:
        got = Fsp.conv_wgrad(fd, dout, pairs[0], pairs[1], pair_num, 9, "bf16", n_dev, n_dev)
    finally:
        Fsp.hints.stop()
    assert int(pair_num.max()) > 256
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))


@pytest.mark.parametrize("rblk", [1, 2, 3, 4])
def test_row_blocked_tiles(cuda_device, rblk):
    """The apply kernel may give one CTA up to four 128-row blocks that share each weight slice.  Force
    every blocking on a problem with a ragged last block (forward, dgrad and bias), bf16 mode."""
    from waveformml_b200 import _lib
    torch.manual_seed(8)
    B = 420
    idx, feats = events(B, 35, 72)
    _lib.load().wfsp_set_option(b"apply_row_blocks", rblk)
    try:
        g, refs = run_pair([spconv.SparseConv2d(72, 40, 3, 1, 1, 1, 1, True)], idx, feats, B, cuda_device, "bf16")
    finally:
        _lib.load().wfsp_set_option(b"apply_row_blocks", 0)
    check_all(g, refs, "bf16", "row blocks %d" % rblk)


def test_fused_stack_takes_bf16_operand_input(cuda_device):
    """Features already in the bf16 operand format (pitch rounded up to 8, e.g. from batcher.pack_batch) are used
    without a cast pass and give the same result as the fp32 features they were rounded from -- on the fused
    path and, through the fallback conversion, on the per-layer path."""
    B = 30
    idx, feats = events(B, 29, 20)  # 20 channels -> pitch 24
    torch.manual_seed(12)
    net = spconv.SparseSequential(spconv.SparseConv2d(20, 16, 3, 1, 1, 1, 1, False), torch.nn.BatchNorm1d(16), torch.nn.ReLU(),
                                  spconv.SubMConv2d(16, 8, 3, bias=False, indice_key="subm0"), spconv.ToDense()).to(cuda_device)
    f32 = feats.to(cuda_device).bfloat16().float()  # values exactly representable in bf16
    f16 = torch.zeros((f32.shape[0], 24), dtype=torch.bfloat16, device=cuda_device)
    f16[:, :20] = f32.bfloat16()
    outs = {}
    for name, fused_on, f in (("fused16", True, f16), ("fused32", True, f32), ("layer16", False, f16)):
        spconv.set_fused(fused_on)
        try:
            net.zero_grad()
            y = net(spconv.SparseConvTensor(f, idx.to(cuda_device), [14, 11], B))
            y.square().sum().backward()
            outs[name] = (y.detach().clone(), net[0].weight.grad.clone())
        finally:
            spconv.set_fused(True)
    torch.testing.assert_close(outs["fused16"][0], outs["fused32"][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(outs["fused16"][1], outs["fused32"][1], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(outs["layer16"][0], outs["fused32"][0], rtol=5e-3, atol=5e-3 * float(outs["fused32"][0].abs().max()))


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 2e-2)])
def test_adjoint_identities_full_size(cuda_device, mode, tol):
    """Size-independent check of dgrad and wgrad at BASELINE's full size (C5: 1024 full-grid events, 157,696
    rows, ~1 M pairs): y = conv(x; W) is linear in x and in W, so for any g
        <y, g> == <x, dL/dx> == <W, dL/dW>      with L = <y, g>.
    Also covers the strided rulebook + inverse convolution pair at that size."""
    B = 1024
    ev = make_events(B, n_samples=1, seed=3, full_grid=True)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(cuda_device)
    torch.manual_seed(1)
    for layers in ([spconv.SparseConv2d(48, 40, 3, 1, 0, 1, 1, False)],
                   [spconv.SparseConv2d(24, 24, 3, 2, 1, 1, 1, False, indice_key="p"),
                    spconv.SparseInverseConv2d(24, 16, 3, "p", bias=False)]):
        net = spconv.SparseSequential(*layers).to(cuda_device)
        for m in net.modules():
            if isinstance(m, spconv.SparseConvolution):
                m.math = mode
        cin = layers[0].in_channels
        x = torch.rand(idx.shape[0], cin, device=cuda_device).requires_grad_(True)
        y = net(spconv.SparseConvTensor(x, idx, [14, 11], B)).features
        g = torch.randn_like(y)
        lhs = (y.detach().double() * g.double()).sum()
        (y * g).sum().backward()
        via_x = (x.detach().double() * x.grad.double()).sum()
        scale = float((y.detach().double().abs() * g.double().abs()).sum())
        assert abs(float(lhs - via_x)) < tol * scale * 1e-2 + tol * abs(float(lhs)), (mode, float(lhs), float(via_x))
        if len(layers) == 1:  # a single layer is also linear in its weight
            w = layers[0].weight
            via_w = (w.detach().double() * w.grad.double()).sum()
            assert abs(float(lhs - via_w)) < tol * scale * 1e-2 + tol * abs(float(lhs)), (mode, float(lhs), float(via_w))
