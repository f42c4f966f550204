"""3-d variant of the rulebook / convolution oracle (net_type "3DConvolution", src/models/SPConvNet.py:42-49:
spatial size [14, 11, n_samples], indices (b, x, y, t)).  Pinned the same way as the 2-d oracle: the dense
identities against torch.nn.functional.conv3d / conv_transpose3d, rulebook invariants, and the embedding
property that a 3-d problem with a singleton last dimension is the 2-d problem."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_cpu as osp
from waveformml_b200.synth import make_events, make_events_3d

SHAPE = [14, 11, 12]


def _voxels(B, seed, cin):
    ev = make_events_3d(B, n_samples=SHAPE[2], seed=seed)
    indices = torch.from_numpy(ev["coords"])[:, [3, 0, 1, 2]].contiguous()
    g = torch.Generator().manual_seed(seed)
    return indices, torch.randn(indices.shape[0], cin, generator=g)


def _dense(indices, feats, B, shape=SHAPE):
    return osp.SparseConvTensor(feats, indices, shape, B).dense()


@pytest.mark.parametrize("k,s,p,d", [(3, 1, 0, 1), (3, 1, 1, 1), (3, 2, 1, 1), (2, 2, 0, 1), ([3, 3, 5], [1, 1, 2], [1, 1, 2], 1),
                                     (3, 1, 2, 2), ([1, 1, 3], 1, 0, 1)])
def test_dense_equivalence_regular_3d(k, s, p, d):
    torch.manual_seed(0)
    B, Cin, Cout = 4, 3, 4
    indices, feats = _voxels(B, 3, Cin)
    feats.requires_grad_(True)
    conv = osp.SparseConv3d(Cin, Cout, k, s, p, d, 1, True)
    y = conv(osp.SparseConvTensor(feats, indices, SHAPE, B))
    X = _dense(indices, feats, B)
    ref = F.conv3d(X, conv.weight.permute(4, 3, 0, 1, 2), None, conv.stride, conv.padding, conv.dilation)
    M = _dense(indices, torch.ones(indices.shape[0], 1), B)
    act = (F.conv3d(M, torch.ones(1, 1, *conv.kernel_size), None, conv.stride, conv.padding, conv.dilation) > 0).float()
    ref = ref + conv.bias.view(1, -1, 1, 1, 1) * act
    yd = y.dense()
    assert yd.shape == ref.shape and y.indices.shape[0] == int(act.sum())
    torch.testing.assert_close(yd, ref, rtol=1e-5, atol=1e-5)
    g = torch.randn_like(ref)
    gF, gW = torch.autograd.grad((yd * g).sum(), [feats, conv.weight], retain_graph=True)
    rF, rW = torch.autograd.grad((ref * g).sum(), [feats, conv.weight])
    torch.testing.assert_close(gF, rF, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gW, rW, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k", [3, [3, 3, 5]])
def test_dense_equivalence_subm_3d(k):
    torch.manual_seed(1)
    B, Cin, Cout = 3, 4, 3
    indices, feats = _voxels(B, 5, Cin)
    feats.requires_grad_(True)
    conv = osp.SubMConv3d(Cin, Cout, k, 2, 9, indice_key="subm0")  # stride / padding arguments are ignored
    y = conv(osp.SparseConvTensor(feats, indices, SHAPE, B))
    X, M = _dense(indices, feats, B), _dense(indices, torch.ones(indices.shape[0], 1), B)
    ref = (F.conv3d(X, conv.weight.permute(4, 3, 0, 1, 2), None, 1, [q // 2 for q in conv.kernel_size])
           + conv.bias.view(1, -1, 1, 1, 1)) * M
    torch.testing.assert_close(y.dense(), ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(y.indices, indices)
    _, pairs, num = x_rb = osp.get_indice_pairs(indices, B, SHAPE, conv.kernel_size, [1] * 3, [q // 2 for q in conv.kernel_size],
                                                [1] * 3, True)
    K = int(np.prod(conv.kernel_size))
    assert num.tolist() == num.flip(0).tolist() and num[K // 2].item() == indices.shape[0]
    g = torch.randn_like(ref)
    gF, gW = torch.autograd.grad((y.dense() * g).sum(), [feats, conv.weight], retain_graph=True)
    rF, rW = torch.autograd.grad((ref * g).sum(), [feats, conv.weight])
    torch.testing.assert_close(gF, rF, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gW, rW, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (2, 2, 0), (3, 2, 1)])
def test_dense_equivalence_inverse_3d(k, s, p):
    torch.manual_seed(2)
    B, C0, C1 = 3, 4, 3
    indices, feats = _voxels(B, 6, C0)
    conv = osp.SparseConv3d(C0, C1, k, s, p, 1, 1, False, indice_key="ind_0")
    inv = osp.SparseInverseConv3d(C1, C1, k, "ind_0", bias=False)
    y = conv(osp.SparseConvTensor(feats, indices, SHAPE, B))
    z = inv(y)
    assert torch.equal(z.indices, indices) and z.spatial_shape == SHAPE
    G = y.dense()
    M = _dense(indices, torch.ones(indices.shape[0], 1), B)
    opad = [o - ((g - 1) * s - 2 * p + k) for o, g in zip(SHAPE, G.shape[2:])]
    ref = F.conv_transpose3d(G, inv.weight.permute(3, 4, 0, 1, 2), None, s, p, output_padding=opad) * M
    torch.testing.assert_close(z.dense(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k,s,p,subm", [(3, 1, 0, False), (3, 2, 1, False), (3, 1, 1, True), (5, 1, 2, True)])
def test_singleton_last_dimension_is_the_2d_rulebook(k, s, p, subm):
    B = 9
    ev = make_events(B, n_samples=1, seed=11)
    i2 = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()
    i3 = torch.cat([i2, torch.zeros(i2.shape[0], 1, dtype=torch.int32)], dim=1).contiguous()
    o2, p2, n2 = osp.get_indice_pairs(i2, B, [14, 11], [k, k], [s, s], [p, p], [1, 1], subm)
    o3, p3, n3 = osp.get_indice_pairs(i3, B, [14, 11, 1], [k, k, 1], [s, s, 1], [p, p, 0], [1, 1, 1], subm)
    assert torch.equal(p2, p3) and torch.equal(n2, n3) and torch.equal(o2, o3[:, :3]) and int(o3[:, 3].abs().sum()) == 0


def test_kat_3d_two_voxels():
    """Hand-derived: voxels a=(0,0,0), b=(0,0,1) in a 2x2x2 grid, k=2, stride 1, padding 1 -> 3x3x3 outputs.
    a (walked first, offsets ascending) creates outputs (1,1,1),(1,1,0),(1,0,1),(1,0,0),(0,1,1),... i.e. the
    candidates are enumerated from the upper bound downwards, offset = sum (in - out*s + p)/d * stride_of_dim."""
    idx = torch.tensor([[0, 0, 0, 0], [0, 0, 0, 1]], dtype=torch.int32)
    outids, pairs, num = osp.get_indice_pairs(idx, 1, [2, 2, 2], [2, 2, 2], [1, 1, 1], [1, 1, 1], [1, 1, 1], False)
    assert num.tolist() == [2] * 8
    # offset 0 = (kx,ky,kz)=(0,0,0): out = in + p = (1,1,1) for a (first row created), (1,1,2) for b
    assert outids[0].tolist() == [0, 1, 1, 1]
    assert outids[pairs[1, 0, 1]].tolist() == [0, 1, 1, 2]
    # offset 1 = kz=1: out z = in z + 1 - 1: a -> (1,1,0) (second row), b -> (1,1,1) = row 0 again
    assert outids[1].tolist() == [0, 1, 1, 0] and pairs[1, 1, :2].tolist() == [1, 0]
    # 8 candidates per voxel, 4 shared cells (z overlap) -> 12 distinct outputs
    assert outids.shape[0] == 12


@pytest.mark.parametrize("k,s,p,d", [([3, 3, 3], [1, 1, 1], [0, 0, 0], [1, 1, 1]), ([3, 2, 5], [2, 1, 2], [1, 0, 2], [1, 1, 1]),
                                     ([2, 3, 3], [1, 1, 1], [1, 2, 0], [1, 2, 3]), ([1, 1, 4], [1, 1, 3], [0, 0, 1], [1, 1, 1])])
def test_rulebook_invariants_3d(k, s, p, d):
    """Every pair (input j, output o, offset (kx,ky,kt)) satisfies in = out*s - p + k*d per dimension; outputs are
    unique and created in first-touch order (walking inputs in order, offsets ascending); the pairs of one offset
    are in ascending input order with no input repeated; counts match the number of valid candidates."""
    B = 5
    indices, _ = _voxels(B, 7, 1)
    outids, pairs, num = osp.get_indice_pairs(indices, B, SHAPE, k, s, p, d, False)
    out_shape = osp.get_conv_output_size(SHAPE, k, s, p, d)
    K = int(np.prod(k))
    inn, out = indices.numpy().astype(np.int64), outids.numpy().astype(np.int64)
    assert len({tuple(r) for r in out.tolist()}) == out.shape[0]
    assert (out[:, 1:] >= 0).all() and (out[:, 1:] < np.array(out_shape)).all()
    first_touch = {}
    total = 0
    for kk in range(K):
        kv = np.unravel_index(kk, k)
        n = int(num[kk])
        total += n
        ji, oi = pairs[0, kk, :n].numpy(), pairs[1, kk, :n].numpy()
        assert (pairs[:, kk, n:] == -1).all()
        assert (np.diff(ji) > 0).all()
        for j, o in zip(ji, oi):
            assert inn[j, 0] == out[o, 0]
            for dim in range(3):
                assert inn[j, 1 + dim] == out[o, 1 + dim] * s[dim] - p[dim] + kv[dim] * d[dim]
            rank = int(j) * K + kk
            first_touch[int(o)] = min(first_touch.get(int(o), rank), rank)
    # brute-force candidate count
    expect = 0
    for row in inn:
        for kk in range(K):
            kv = np.unravel_index(kk, k)
            ok = True
            for dim in range(3):
                num_ = row[1 + dim] + p[dim] - kv[dim] * d[dim]
                if num_ < 0 or num_ % s[dim] or num_ // s[dim] >= out_shape[dim]:
                    ok = False
            expect += ok
    assert total == expect
    order = [first_touch[o] for o in range(out.shape[0])]
    assert order == sorted(order)  # output rows are numbered in the order they are first touched
