"""Pins the CPU oracle (oracle/) against the only anchors that exist for this path:
SURVEY.md A.6 hand-derived known-answer rulebooks and the A.5 dense-convolution identities
(the same idea as the reference's src/models/DenseConvNet.py:26-34).  The reference itself
ships no tests or golden vectors for its spconv calls (parity unpinned, see oracle header)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_cpu as osp
from waveformml_b200.synth import make_events


def idx(rows):
    return torch.tensor(rows, dtype=torch.int32).reshape(-1, 3)


def rb(rows, k, s=1, p=0, d=1, subm=False, shape=(14, 11), B=1):
    return osp.get_indice_pairs(idx(rows), B, list(shape), [k, k], [s, s], [p, p], [d, d], subm)


def test_kat1_regular_k3():
    outids, pairs, num = rb([(0, 5, 5), (0, 5, 6)], 3)
    exp = [(5, 5), (5, 4), (5, 3), (4, 5), (4, 4), (4, 3), (3, 5), (3, 4), (3, 3), (5, 6), (4, 6), (3, 6)]
    assert outids.tolist() == [[0, x, y] for x, y in exp]
    assert num.tolist() == [2] * 9
    assert pairs[0, :, :2].tolist() == [[0, 1]] * 9
    assert pairs[1, :, :2].tolist() == [[0, 9], [1, 0], [2, 1], [3, 10], [4, 3], [5, 4], [6, 11], [7, 6], [8, 7]]
    assert osp.get_conv_output_size([14, 11], [3, 3], [1, 1], [0, 0], [1, 1]) == [12, 9]


def test_kat2_subm_k3():
    outids, pairs, num = rb([(0, 5, 5), (0, 5, 6), (0, 7, 7)], 3, subm=True)
    assert num.tolist() == [0, 0, 0, 1, 3, 1, 0, 0, 0]
    assert (pairs[0, 3, 0].item(), pairs[1, 3, 0].item()) == (0, 1)
    assert pairs[0, 4, :3].tolist() == [0, 1, 2] and pairs[1, 4, :3].tolist() == [0, 1, 2]
    assert (pairs[0, 5, 0].item(), pairs[1, 5, 0].item()) == (1, 0)
    assert outids.tolist() == [[0, 5, 5], [0, 5, 6], [0, 7, 7]]


def test_kat3_strided():
    outids, pairs, num = rb([(0, 2, 2)], 3, s=2)
    assert osp.get_conv_output_size([14, 11], [3, 3], [2, 2], [0, 0], [1, 1]) == [6, 5]
    assert outids.tolist() == [[0, 1, 1], [0, 1, 0], [0, 0, 1], [0, 0, 0]]
    assert num.tolist() == [1, 0, 1, 0, 0, 0, 1, 0, 1]
    assert [pairs[1, k, 0].item() for k in (0, 2, 6, 8)] == [0, 1, 2, 3]


def test_kat4_even_kernel():
    outids, pairs, num = rb([(0, 0, 0)], 2)
    assert outids.tolist() == [[0, 0, 0]] and num.tolist() == [1, 0, 0, 0]
    outids, pairs, num = rb([(0, 13, 10)], 2)
    assert outids.tolist() == [[0, 12, 9]] and num.tolist() == [0, 0, 0, 1]


def _events(B, seed, C, full=False):
    ev = make_events(B, n_samples=1, seed=seed, full_grid=full)
    c = torch.from_numpy(ev["coords"])
    indices = c[:, [2, 0, 1]].contiguous()
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(indices.shape[0], C, generator=g)
    return indices, feats


@pytest.mark.parametrize("k,s,p,d", [(3, 1, 0, 1), (3, 1, 1, 1), (3, 2, 1, 1), (2, 1, 0, 1), (5, 1, 2, 1),
                                     (3, 1, 2, 2), (2, 2, 0, 1), (3, 3, 0, 1), (5, 2, 0, 1)])
def test_rulebook_properties(k, s, p, d):
    B = 7
    indices, _ = _events(B, 5, 1)
    outids, pairs, num = osp.get_indice_pairs(indices, B, [14, 11], [k, k], [s, s], [p, p], [d, d])
    N = indices.shape[0]
    # outputs unique, inside the output grid
    oshape = osp.get_conv_output_size([14, 11], [k, k], [s, s], [p, p], [d, d])
    assert len({tuple(r) for r in outids.tolist()}) == outids.shape[0]
    assert (outids[:, 1] >= 0).all() and (outids[:, 1] < oshape[0]).all()
    assert (outids[:, 2] >= 0).all() and (outids[:, 2] < oshape[1]).all()
    first_seen = []
    for kk in range(k * k):
        n = num[kk].item()
        kx, ky = divmod(kk, k)
        i_rows, o_rows = pairs[0, kk, :n].long(), pairs[1, kk, :n].long()
        assert (pairs[:, kk, n:] == -1).all()
        assert (i_rows[1:] > i_rows[:-1]).all()  # ascending input order, each input once per offset
        ii, oo = indices[i_rows], outids[o_rows]
        assert (ii[:, 0] == oo[:, 0]).all()
        assert (ii[:, 1] == oo[:, 1] * s - p + kx * d).all()
        assert (ii[:, 2] == oo[:, 2] * s - p + ky * d).all()
    # first-touch order: walking inputs in order and offsets ascending reproduces outids order
    by_in = {}
    for kk in range(k * k):
        n = num[kk].item()
        for i, o in zip(pairs[0, kk, :n].tolist(), pairs[1, kk, :n].tolist()):
            by_in.setdefault(i, []).append((kk, o))
    seen = set()
    for i in range(N):
        for kk, o in sorted(by_in.get(i, [])):
            if o not in seen:
                seen.add(o)
                first_seen.append(o)
    assert first_seen == list(range(outids.shape[0]))


@pytest.mark.parametrize("k", [3, 5, 15])
def test_subm_properties(k):
    B = 9
    indices, _ = _events(B, 11, 1)
    outids, pairs, num = osp.get_indice_pairs(indices, B, [14, 11], [k, k], [1, 1], [k // 2] * 2, [1, 1], subm=True)
    K = k * k
    assert num.tolist() == num.flip(0).tolist()  # symmetric
    assert num[K // 2].item() == indices.shape[0]
    n = num[K // 2].item()
    assert pairs[0, K // 2, :n].tolist() == list(range(n)) == pairs[1, K // 2, :n].tolist()
    assert int(np.argmax(num.numpy())) == K // 2


def _dense(indices, feats, B, shape):
    return osp.SparseConvTensor(feats, indices, shape, B).dense()


@pytest.mark.parametrize("k,s,p,d,bias", [(3, 1, 0, 1, False), (3, 1, 1, 1, True), (3, 2, 1, 1, False),
                                          (2, 1, 0, 1, False), (5, 2, 2, 1, True), (3, 1, 2, 2, False), (1, 1, 0, 1, True)])
def test_dense_equivalence_regular(k, s, p, d, bias):
    torch.manual_seed(0)
    B, Cin, Cout = 6, 5, 4
    indices, feats = _events(B, 3, Cin)
    feats = feats.double().float().requires_grad_(True)
    conv = osp.SparseConv2d(Cin, Cout, k, s, p, d, 1, bias)
    x = osp.SparseConvTensor(feats, indices, [14, 11], B)
    y = conv(x)
    yd = y.dense()
    X = _dense(indices, feats, B, [14, 11])
    Wt = conv.weight.permute(3, 2, 0, 1)
    ref = F.conv2d(X, Wt, None, s, p if k > 1 else 0, d) if k > 1 else F.conv2d(X, Wt)
    if bias:
        if k > 1:
            M = _dense(indices, torch.ones(indices.shape[0], 1), B, [14, 11])
            act = (F.conv2d(M, torch.ones(1, 1, k, k), None, s, p, d) > 0).float()
        else:
            act = _dense(indices, torch.ones(indices.shape[0], 1), B, [14, 11])
        ref = ref + conv.bias.view(1, -1, 1, 1) * act
    assert yd.shape == ref.shape
    torch.testing.assert_close(yd, ref, rtol=1e-5, atol=1e-5)
    # gradients through the oracle's hand-written backward == torch autograd through conv2d
    g = torch.randn_like(ref)
    gF, gW = torch.autograd.grad((yd * g).sum(), [feats, conv.weight], retain_graph=True)
    rF, rW = torch.autograd.grad((ref * g).sum(), [feats, conv.weight])
    torch.testing.assert_close(gF, rF, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gW, rW, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k", [3, 5])
def test_dense_equivalence_subm(k):
    torch.manual_seed(1)
    B, Cin, Cout = 5, 6, 3
    indices, feats = _events(B, 4, Cin)
    feats.requires_grad_(True)
    conv = osp.SubMConv2d(Cin, Cout, k, 1, 7, indice_key="subm0")  # padding arg is ignored
    y = conv(osp.SparseConvTensor(feats, indices, [14, 11], B))
    X = _dense(indices, feats, B, [14, 11])
    M = _dense(indices, torch.ones(indices.shape[0], 1), B, [14, 11])
    ref = (F.conv2d(X, conv.weight.permute(3, 2, 0, 1), None, 1, k // 2) + conv.bias.view(1, -1, 1, 1)) * M
    torch.testing.assert_close(y.dense(), ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(y.indices, indices)
    g = torch.randn_like(ref)
    gF, gW = torch.autograd.grad((y.dense() * g).sum(), [feats, conv.weight], retain_graph=True)
    rF, rW = torch.autograd.grad((ref * g).sum(), [feats, conv.weight])
    torch.testing.assert_close(gF, rF, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gW, rW, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (2, 1, 0), (3, 2, 1)])
def test_dense_equivalence_inverse(k, s, p):
    torch.manual_seed(2)
    B, C0, C1 = 4, 5, 3
    indices, feats = _events(B, 6, C0)
    conv = osp.SparseConv2d(C0, C1, k, s, p, 1, 1, False, indice_key="ind_0")
    inv = osp.SparseInverseConv2d(C1, C1, k, "ind_0", bias=False)
    x = osp.SparseConvTensor(feats, indices, [14, 11], B)
    y = conv(x)
    y.features = y.features.detach().requires_grad_(True)
    z = inv(y)
    assert torch.equal(z.indices, indices) and z.spatial_shape == [14, 11]
    G = y.dense()
    M = _dense(indices, torch.ones(indices.shape[0], 1), B, [14, 11])
    opad = [o - ((g - 1) * s - 2 * p + k) for o, g in zip((14, 11), G.shape[2:])]
    ref = F.conv_transpose2d(G, inv.weight.permute(2, 3, 0, 1), None, s, p, output_padding=opad) * M
    torch.testing.assert_close(z.dense(), ref, rtol=1e-5, atol=1e-5)
    g = torch.randn_like(ref)
    gF, gW = torch.autograd.grad((z.dense() * g).sum(), [y.features, inv.weight], retain_graph=True)
    rF_dense, rW = torch.autograd.grad((ref * g).sum(), [G, inv.weight])
    oi = y.indices.long()
    torch.testing.assert_close(gF, rF_dense[oi[:, 0], :, oi[:, 1], oi[:, 2]], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gW, rW, rtol=1e-5, atol=1e-5)


def test_to_dense_c_matches_torch():
    indices, feats = _events(5, 9, 7)
    a = osp.SparseConvTensor(feats, indices, [14, 11], 5).dense()
    b = osp.to_dense_c(feats, indices, 5, [14, 11])
    assert torch.equal(a, b)


def test_empty_and_single():
    outids, pairs, num = rb([], 3)
    assert outids.shape == (0, 3) and num.tolist() == [0] * 9 and pairs.shape == (2, 9, 0)
    outids, pairs, num = rb([(0, 0, 0)], 3, p=1)
    assert num.sum().item() == 4  # corner: 4 of 9 candidates inside the grid
