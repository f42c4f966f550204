"""SparseConvNet-signature facade (src/models/SCNet.py:62-77,
config/examples/OPs3ns_SCNet.json:22-66): the stack the reference builds with `sparseconvnet`, checked
against dense convolutions of the densified input (SURVEY.md A.5 identities)."""
import pytest
import torch
import torch.nn.functional as F

import sparseconvnet as scn
from waveformml_b200 import spconv
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu


def _batch(B, C, seed, dev):
    ev = make_events(B, n_samples=1, seed=seed)
    coords = torch.from_numpy(ev["coords"]).long()  # (x, y, batch): the order SCNet.forward hands to InputLayer
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(coords.shape[0], C, generator=g)
    dense = torch.zeros(B, C, 14, 11)
    dense[coords[:, 2], :, coords[:, 0], coords[:, 1]] = feats
    return coords.to(dev), feats.to(dev), dense


def _w(conv):  # [volume, nIn, nOut] -> conv2d weight [nOut, nIn, kH, kW]
    k = conv.filter_size
    return conv.weight.detach().cpu().view(k[0], k[1], conv.nIn, conv.nOut).permute(3, 2, 0, 1).contiguous()


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_ops3ns_stack_matches_dense_convolutions(cuda_device, mode, tol):
    """Convolution(2,300,37,1,1) . Convolution(2,37,37,3,1) . Convolution(2,37,18,3,2) . SparseToDense(2,18)"""
    torch.manual_seed(1)
    B = 21
    coords, feats, dense = _batch(B, 300, 5, cuda_device)
    convs = [scn.Convolution(2, 300, 37, 1, 1, False), scn.Convolution(2, 37, 37, 3, 1, False),
             scn.Convolution(2, 37, 18, 3, 2, False)]
    net = scn.Sequential(*convs, scn.SparseToDense(2, 18)).to(cuda_device)
    spconv.set_math_mode(mode)
    try:
        x = scn.InputLayer(2, torch.LongTensor([14, 11]), mode=0)([coords, feats])
        y = net(x)
    finally:
        spconv.set_math_mode("bf16")
    ref = dense
    for c, s in zip(convs, (1, 1, 2)):
        ref = F.conv2d(ref, _w(c), None, stride=s)
    assert tuple(y.shape) == tuple(ref.shape) == (B, 18, 5, 4)
    err = (y.detach().cpu() - ref).abs().max() / ref.abs().max()
    assert float(err) < tol, float(err)


def test_submanifold_and_batchnormrelu(cuda_device):
    torch.manual_seed(2)
    B = 17
    coords, feats, dense = _batch(B, 12, 9, cuda_device)
    conv = scn.SubmanifoldConvolution(2, 12, 10, 3, True)
    bn = scn.BatchNormReLU(10)
    net = scn.Sequential(conv, bn, scn.OutputLayer(2)).to(cuda_device)
    spconv.set_math_mode("fp32")
    try:
        out = net(scn.InputLayer(2, [14, 11], mode=0)([coords, feats]))
    finally:
        spconv.set_math_mode("bf16")
    c = coords.cpu()
    full = F.conv2d(dense, _w(conv), conv.bias.detach().cpu(), padding=1)
    rows = full[c[:, 2], :, c[:, 0], c[:, 1]]                      # submanifold: outputs at the input sites, same order
    ref_bn = torch.nn.BatchNorm1d(10, eps=1e-4, momentum=0.1)      # SparseConvNet momentum 0.9 == torch momentum 0.1
    ref = torch.relu(ref_bn(rows))
    torch.testing.assert_close(out.detach().cpu(), ref, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(bn.running_mean.cpu(), ref_bn.running_mean, rtol=1e-4, atol=1e-6)


def test_fused_and_per_layer_agree_with_facade_bn(cuda_device):
    """BatchNormReLU carries its own ReLU: both execution paths must apply it exactly once."""
    B = 30
    coords, feats, _ = _batch(B, 20, 3, cuda_device)
    outs = []
    for fused_on in (True, False):
        torch.manual_seed(4)
        net = scn.Sequential(scn.Convolution(2, 20, 16, 3, 1, False), scn.BatchNormReLU(16),
                             scn.SubmanifoldConvolution(2, 16, 8, 3, False), scn.SparseToDense(2, 8)).to(cuda_device)
        spconv.set_fused(fused_on)
        try:
            y = net(scn.InputLayer(2, [14, 11], mode=0)([coords, feats]))
        finally:
            spconv.set_fused(True)
        assert float(y.detach().min()) < 0  # the last layer has no activation
        outs.append(y.detach())
    torch.testing.assert_close(outs[0], outs[1], rtol=5e-3, atol=5e-3 * float(outs[1].abs().max()))


def test_dimension_3_matches_dense_convolutions(cuda_device):
    """SCNet.py:53-55 (net_type "3DConvolution"): coords (x, y, t, batch), spatial size [14, 11, n_samples]."""
    from waveformml_b200.synth import make_events_3d
    torch.manual_seed(6)
    B, T, C = 9, 10, 2
    ev = make_events_3d(B, n_samples=T, seed=12)
    coords = torch.from_numpy(ev["coords"]).long()
    feats = torch.rand(coords.shape[0], C)
    dense = torch.zeros(B, C, 14, 11, T)
    dense[coords[:, 3], :, coords[:, 0], coords[:, 1], coords[:, 2]] = feats
    sub = scn.SubmanifoldConvolution(3, C, 8, 3, False)
    conv = scn.Convolution(3, 8, 6, [3, 3, 2], [1, 1, 2], False)
    net = scn.Sequential(sub, conv, scn.SparseToDense(3, 6)).to(cuda_device)
    spconv.set_math_mode("fp32")
    try:
        y = net(scn.InputLayer(3, [14, 11, T], mode=0)([coords.to(cuda_device), feats.to(cuda_device)]))
    finally:
        spconv.set_math_mode("bf16")

    def w3(c):
        k = c.filter_size
        return c.weight.detach().cpu().view(k[0], k[1], k[2], c.nIn, c.nOut).permute(4, 3, 0, 1, 2).contiguous()
    mask = (dense.abs().sum(1, keepdim=True) > 0).float()
    ref = F.conv3d(F.conv3d(dense, w3(sub), None, padding=1) * mask, w3(conv), None, stride=[1, 1, 2])
    assert tuple(y.shape) == tuple(ref.shape) == (B, 6, 12, 9, 5)
    torch.testing.assert_close(y.detach().cpu(), ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("fused_on", [True, False])
def test_submanifold_layers_either_side_of_a_stride(cuda_device, fused_on):
    """The canonical SparseConvNet pattern SubmanifoldConvolution -> Convolution(stride 2) -> SubmanifoldConvolution
    with ONE filter size: the second submanifold layer must get the rulebook of the strided active set, not the
    cached one of the input resolution (shared `scn_subm3x3` key)."""
    torch.manual_seed(8)
    B = 19
    coords, feats, dense = _batch(B, 6, 21, cuda_device)
    s1 = scn.SubmanifoldConvolution(2, 6, 8, 3, False)
    cv = scn.Convolution(2, 8, 8, 2, 2, False)
    s2 = scn.SubmanifoldConvolution(2, 8, 5, 3, False)
    net = scn.Sequential(s1, cv, s2, scn.SparseToDense(2, 5)).to(cuda_device)
    spconv.set_math_mode("fp32")
    spconv.set_fused(fused_on)
    try:
        y = net(scn.InputLayer(2, [14, 11], mode=0)([coords, feats]))
    finally:
        spconv.set_math_mode("bf16")
        spconv.set_fused(True)
    m0 = (dense.abs().sum(1, keepdim=True) > 0).float()
    a = F.conv2d(dense, _w(s1), None, padding=1) * m0
    b = F.conv2d(a, _w(cv), None, stride=2)
    m1 = (F.conv2d(m0, torch.ones(1, 1, 2, 2), None, stride=2) > 0).float()  # active iff the receptive field holds an input
    ref = F.conv2d(b * m1, _w(s2), None, padding=1) * m1
    assert tuple(y.shape) == tuple(ref.shape) == (B, 5, 7, 5)
    torch.testing.assert_close(y.detach().cpu(), ref, rtol=1e-4, atol=1e-5)


def test_shared_indice_key_across_resolutions_rebuilds(cuda_device):
    """spconv layers: SubMConv2d('k') -> SparseConv2d(stride 2) -> SubMConv2d('k').  Upstream would reuse the first
    rulebook for the third layer (wrong rows); the cache entry is validated against the input and rebuilt."""
    import spconv as sp
    torch.manual_seed(9)
    B = 11
    coords, feats, dense = _batch(B, 4, 33, cuda_device)
    idx = coords[:, [2, 0, 1]].to(torch.int32).contiguous()
    l1 = sp.SubMConv2d(4, 6, 3, bias=False, indice_key="k").to(cuda_device)
    l2 = sp.SparseConv2d(6, 6, 3, 2, 1, bias=False).to(cuda_device)
    l3 = sp.SubMConv2d(6, 3, 3, bias=False, indice_key="k").to(cuda_device)
    sp.set_math_mode("fp32")
    try:
        x = sp.SparseConvTensor(feats, idx, [14, 11], B)
        y = l3(l2(l1(x))).dense()
    finally:
        sp.set_math_mode("bf16")

    def w(c):
        return c.weight.detach().cpu().permute(3, 2, 0, 1).contiguous()
    m0 = (dense.abs().sum(1, keepdim=True) > 0).float()
    a = F.conv2d(dense, w(l1), None, padding=1) * m0
    b = F.conv2d(a, w(l2), None, stride=2, padding=1)
    m1 = (F.conv2d(m0, torch.ones(1, 1, 3, 3), None, stride=2, padding=1) > 0).float()
    ref = F.conv2d(b * m1, w(l3), None, padding=1) * m1
    torch.testing.assert_close(y.detach().cpu(), ref, rtol=1e-4, atol=1e-5)
