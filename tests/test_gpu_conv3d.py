"""3-d sparse convolutions (net_type "3DConvolution", src/models/SPConvNet.py:42-49; spconv.SparseConv3d /
SubMConv3d, src/utils/ModelValidation.py:24-31) through the C-ABI vs the CPU oracle: rulebooks bit-exact,
features / gradients within the fp32 and bf16 tolerances of SURVEY.md 8d."""
import pytest
import torch
from torch import nn

from oracle import spconv_cpu as osp
from waveformml_b200 import _lib, spconv
from waveformml_b200.spconv import functional as Fsp
from waveformml_b200.spconv import ops
from waveformml_b200.synth import make_events, make_events_3d

pytestmark = pytest.mark.gpu


def _voxels(B, T, seed):
    ev = make_events_3d(B, n_samples=T, seed=seed)
    return torch.from_numpy(ev["coords"])[:, [3, 0, 1, 2]].contiguous()


def _same_rulebook(g, e_out, e_pairs, e_num, n):
    n_out = e_out.shape[0]
    assert torch.equal(g.pair_num.cpu(), e_num)
    assert torch.equal(g.outids.cpu()[:n_out], e_out)
    assert torch.equal(g.pairs.cpu()[:, :, :n], e_pairs)


@pytest.mark.parametrize("B,T", [(5, 8), (64, 16), (300, 24)])  # one-CTA builder, phase kernels (direct table)
@pytest.mark.parametrize("k,s,p,d,subm", [(3, 1, 0, 1, False), (3, 2, 1, 1, False), ([3, 3, 5], [1, 1, 2], [1, 1, 2], 1, False),
                                          (2, 2, 0, 1, False), (3, 1, 2, 2, False), (3, 1, 1, 1, True), ([3, 3, 5], 1, 0, 1, True),
                                          (3, 1, 0, 2, True)])
def test_rulebook_3d_matches_oracle(cuda_device, B, T, k, s, p, d, subm):
    idx = _voxels(B, T, 21)
    shape = [14, 11, T]
    ks, st, pd, dl = (ops._listn(v, 3) for v in (k, s, p, d))
    if subm:
        st, pd = [1, 1, 1], [q // 2 for q in ks]
    e_out, e_pairs, e_num = osp.get_indice_pairs(idx, B, shape, ks, st, pd, dl, subm)
    g = ops.build_rulebook(idx.to(cuda_device), B, shape, ks, st, pd, dl, subm, check_duplicates=True)
    _same_rulebook(g, e_out, e_pairs, e_num, idx.shape[0])
    assert g.outids.shape == e_out.shape


def test_rulebook_3d_hash_table_and_capacity_path(cuda_device):
    B, T = 40, 16
    idx = _voxels(B, T, 3)
    n = idx.shape[0]
    shape, ks = [14, 11, T], [3, 3, 3]
    e_out, e_pairs, e_num = osp.get_indice_pairs(idx, B, shape, ks, [2, 2, 2], [1, 1, 1], [1, 1, 1], False)
    lib = _lib.load()
    lib.wfsp_set_option(b"rulebook_force_hash", 1)
    try:
        g = ops.build_rulebook(idx.to(cuda_device), B, shape, ks, [2, 2, 2], [1, 1, 1], [1, 1, 1], False)
    finally:
        lib.wfsp_set_option(b"rulebook_force_hash", 0)
    _same_rulebook(g, e_out, e_pairs, e_num, n)
    # graph path: capacity-sized buffer, live count on the device
    padded = torch.zeros((n + 500, 4), dtype=torch.int32, device=cuda_device)
    padded[:n] = idx.to(cuda_device)
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    g = ops.build_rulebook(padded, B, shape, ks, [2, 2, 2], [1, 1, 1], [1, 1, 1], False, n_rows=n_dev)
    assert int(g.n_out_dev.item()) == e_out.shape[0]
    _same_rulebook(g, e_out, e_pairs, e_num, n)


def test_singleton_depth_equals_2d(cuda_device):
    B = 31
    ev = make_events(B, n_samples=1, seed=8)
    i2 = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(cuda_device)
    i3 = torch.cat([i2, torch.zeros_like(i2[:, :1])], dim=1).contiguous()
    for subm in (False, True):
        a = ops.build_rulebook(i2, B, [14, 11], [3, 3], [1, 1], [1, 1], [1, 1], subm)
        b = ops.build_rulebook(i3, B, [14, 11, 1], [3, 3, 1], [1, 1, 1], [1, 1, 0], [1, 1, 1], subm)
        assert torch.equal(a.pairs, b.pairs) and torch.equal(a.pair_num, b.pair_num)
        assert torch.equal(a.outids, b.outids[:, :3]) and torch.equal(a.nbr_out, b.nbr_out)


def _copy_params(dst, src):
    with torch.no_grad():
        for pd_, ps_ in zip(dst.parameters(), src.parameters()):
            pd_.copy_(ps_)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv3d_stack_forward_backward(cuda_device, mode):
    """SubMConv3d -> SparseConv3d (strided) -> SparseInverseConv3d -> 1x1x1 -> ToDense, against the oracle's
    gather-mm-scatter restatement (bf16: operands rounded as the tensor-core path rounds them)."""
    torch.manual_seed(0)
    B, T, C = 12, 10, 2
    idx = _voxels(B, T, 17)
    feats = torch.randn(idx.shape[0], C)
    shape = [14, 11, T]

    def build(m):
        return m.SparseSequential(
            m.SubMConv3d(C, 24, 3, indice_key="subm0"), nn.ReLU(),
            m.SparseConv3d(24, 40, [3, 3, 2], [1, 1, 2], [1, 1, 0], indice_key="down"), nn.ReLU(),
            m.SparseInverseConv3d(40, 16, [3, 3, 2], "down", bias=False),
            m.SparseConv3d(16, 8, 1),
            m.ToDense())
    ref_net, net = build(osp), build(spconv)
    _copy_params(net, ref_net)
    net = net.to(cuda_device)
    osp.set_operand_rounding("bf16" if mode == "bf16" else None)
    try:
        fr = feats.clone().requires_grad_(True)
        yr = ref_net(osp.SparseConvTensor(fr, idx, shape, B))
        gy = torch.randn_like(yr)
        (yr * gy).sum().backward()
    finally:
        osp.set_operand_rounding(None)
    spconv.set_math_mode(mode)
    try:
        fg = feats.to(cuda_device).requires_grad_(True)
        y = net(spconv.SparseConvTensor(fg, idx.to(cuda_device), shape, B))
        (y * gy.to(cuda_device)).sum().backward()
    finally:
        spconv.set_math_mode("bf16")
    assert y.shape == yr.shape == (B, 8, 14, 11, T)
    rtol, atol = (2e-3, 1e-4) if mode == "fp32" else (2e-2, 2e-2)
    torch.testing.assert_close(y.cpu(), yr, rtol=rtol, atol=atol * float(yr.abs().max()))
    torch.testing.assert_close(fg.grad.cpu(), fr.grad, rtol=rtol, atol=atol * float(fr.grad.abs().max()))
    for pg, pr in zip(net.parameters(), ref_net.parameters()):
        torch.testing.assert_close(pg.grad.cpu(), pr.grad, rtol=rtol, atol=atol * float(pr.grad.abs().max()))
