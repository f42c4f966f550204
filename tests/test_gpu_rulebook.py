"""GPU rulebook builder vs the CPU oracle: bit-exact (torch.equal) output rows, pairs [2,K,N]
including the -1 padding, and pair counts; both table modes (direct grid / open-addressing hash);
plus the neighbour tables derived from the pairs."""
import numpy as np
import pytest
import torch

from oracle import spconv_cpu as osp
from waveformml_b200 import _lib
from waveformml_b200.spconv import ops
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu

GEOMS = [(3, 1, 0, 1), (3, 1, 1, 1), (3, 2, 1, 1), (2, 1, 0, 1), (5, 1, 2, 1), (3, 1, 2, 2), (2, 2, 0, 1),
         (3, 3, 0, 1), (5, 2, 0, 1), (15, 1, 7, 1), (4, 1, 1, 1), (7, 1, 0, 1)]


def _indices(B, seed, full=False):
    ev = make_events(B, n_samples=1, seed=seed, full_grid=full)
    return torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()


def _check(indices, B, k, s, p, d, subm, dev):
    if subm:
        ref = osp.get_indice_pairs(indices, B, [14, 11], [k, k], [1, 1], [k // 2] * 2, [d, d], True)
        rb = ops.build_rulebook(indices.to(dev), B, [14, 11], [k, k], [1, 1], [k // 2] * 2, [d, d], True)
    else:
        ref = osp.get_indice_pairs(indices, B, [14, 11], [k, k], [s, s], [p, p], [d, d], False)
        rb = ops.build_rulebook(indices.to(dev), B, [14, 11], [k, k], [s, s], [p, p], [d, d], False)
    outids, pairs, num = ref
    assert torch.equal(rb.pair_num.cpu(), num), (k, s, p, d, subm)
    assert torch.equal(rb.outids.cpu(), outids), (k, s, p, d, subm)
    assert torch.equal(rb.pairs.cpu(), pairs), (k, s, p, d, subm)
    assert int(rb.dup_flag.item()) == 0
    # neighbour tables are consistent with the pairs
    K, N = pairs.shape[1], pairs.shape[2]
    nbr_out = torch.full((outids.shape[0], K), -1, dtype=torch.int32)
    nbr_in = torch.full((N, K), -1, dtype=torch.int32)
    for kk in range(K):
        n = int(num[kk])
        i, o = pairs[0, kk, :n].long(), pairs[1, kk, :n].long()
        nbr_out[o, kk] = i.int()
        nbr_in[i, kk] = o.int()
    assert torch.equal(rb.nbr_out.cpu(), nbr_out) and torch.equal(rb.nbr_in.cpu(), nbr_in)
    return rb


@pytest.mark.parametrize("force_hash", [0, 1])
@pytest.mark.parametrize("k,s,p,d", GEOMS)
def test_regular_matches_oracle(cuda_device, k, s, p, d, force_hash):
    _lib.load().wfsp_set_option(b"rulebook_force_hash", force_hash)
    try:
        _check(_indices(37, 5), 37, k, s, p, d, False, cuda_device)
    finally:
        _lib.load().wfsp_set_option(b"rulebook_force_hash", 0)


@pytest.mark.parametrize("force_hash", [0, 1])
@pytest.mark.parametrize("k,d", [(3, 1), (5, 1), (9, 1), (15, 1), (3, 2)])
def test_subm_matches_oracle(cuda_device, k, d, force_hash):
    _lib.load().wfsp_set_option(b"rulebook_force_hash", force_hash)
    try:
        _check(_indices(41, 6), 41, k, 1, 0, d, True, cuda_device)
    finally:
        _lib.load().wfsp_set_option(b"rulebook_force_hash", 0)


def test_kats(cuda_device):
    t = lambda rows: torch.tensor(rows, dtype=torch.int32).reshape(-1, 3)
    rb = _check(t([(0, 5, 5), (0, 5, 6)]), 1, 3, 1, 0, 1, False, cuda_device)  # KAT-1
    assert rb.outids.cpu()[:, 1:].tolist() == [[5, 5], [5, 4], [5, 3], [4, 5], [4, 4], [4, 3], [3, 5], [3, 4],
                                               [3, 3], [5, 6], [4, 6], [3, 6]]
    rb = _check(t([(0, 5, 5), (0, 5, 6), (0, 7, 7)]), 1, 3, 1, 0, 1, True, cuda_device)  # KAT-2
    assert rb.pair_num.cpu().tolist() == [0, 0, 0, 1, 3, 1, 0, 0, 0]
    rb = _check(t([(0, 2, 2)]), 1, 3, 2, 0, 1, False, cuda_device)  # KAT-3
    assert rb.outids.cpu().tolist() == [[0, 1, 1], [0, 1, 0], [0, 0, 1], [0, 0, 0]]
    rb = _check(t([(0, 13, 10)]), 1, 2, 1, 0, 1, False, cuda_device)  # KAT-4
    assert rb.outids.cpu().tolist() == [[0, 12, 9]] and rb.pair_num.cpu().tolist() == [0, 0, 0, 1]


def test_empty_and_ragged(cuda_device):
    e = torch.zeros((0, 3), dtype=torch.int32)
    rb = _check(e, 4, 3, 1, 0, 1, False, cuda_device)
    assert rb.outids.shape == (0, 3) and rb.pairs.shape == (2, 9, 0)
    _check(e, 4, 3, 1, 0, 1, True, cuda_device)
    # events with no hits in the middle of the batch, single-hit events, multi-block inputs
    idx = _indices(700, 8)
    keep = (idx[:, 0] % 3) != 1
    _check(idx[keep].contiguous(), 700, 3, 1, 0, 1, False, cuda_device)
    _check(idx[keep].contiguous(), 700, 3, 1, 0, 1, True, cuda_device)


def test_c1_and_chained_layers(cuda_device):
    """C1 batch (64 events): layer-2 and layer-3 rulebooks of the GEP stack, chained."""
    idx = _indices(64, 1234)
    rb1 = _check(idx, 64, 3, 1, 0, 1, False, cuda_device)
    rb2 = _check(rb1.outids.cpu().contiguous(), 64, 3, 1, 0, 1, False, cuda_device)
    assert rb2.outids.shape[0] > rb1.outids.shape[0] > idx.shape[0]


def test_mid_sizes_both_builders(cuda_device):
    """~2,700 rows: the single-launch builder walks them in three rounds of 1024; its output (~9,900 rows)
    is past that builder's limit and goes through the multi-kernel phases with the direct table."""
    idx = _indices(900, 77)
    assert 2048 < idx.shape[0] <= 8192
    rb1 = _check(idx, 900, 3, 1, 0, 1, False, cuda_device)
    _check(idx, 900, 5, 1, 0, 1, True, cuda_device)
    assert rb1.outids.shape[0] > 8192
    _check(rb1.outids.cpu().contiguous(), 900, 3, 1, 0, 1, False, cuda_device)
    _check(rb1.outids.cpu().contiguous(), 900, 3, 1, 0, 1, True, cuda_device)


def test_full_grid_1024_matches_oracle(cuda_device):
    """C5 at full size: 1024 events x 154 cells = 157,696 rows, ~1 M pairs."""
    idx = _indices(1024, 1, full=True)
    rb = _check(idx, 1024, 3, 1, 0, 1, False, cuda_device)
    assert rb.outids.shape[0] == 1024 * 12 * 9 and int(rb.pair_num.sum()) == 995328
    rb = _check(idx, 1024, 3, 1, 0, 1, True, cuda_device)
    assert int(rb.pair_num.sum()) == 1269760


def test_unsorted_input_first_touch_order(cuda_device):
    idx = _indices(23, 3)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(idx.shape[0]))
    _check(idx[perm].contiguous(), 23, 3, 1, 1, 1, False, cuda_device)
    _check(idx[perm].contiguous(), 23, 3, 1, 1, 1, True, cuda_device)


def test_duplicate_coordinates_flagged(cuda_device):
    idx = torch.tensor([(0, 5, 5), (0, 5, 5), (0, 6, 6)], dtype=torch.int32)
    ref = osp.get_indice_pairs(idx, 1, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False)
    rb = ops.build_rulebook(idx.to(cuda_device), 1, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False,
                            check_duplicates=False)
    assert torch.equal(rb.pairs.cpu(), ref[1]) and torch.equal(rb.outids.cpu(), ref[0])  # the rulebook itself is exact
    assert int(rb.dup_flag.item()) == 1
    # default: the eager path raises (upstream would SUM the duplicates; silently keeping one is not acceptable)
    with pytest.raises(RuntimeError, match="duplicate"):
        ops.build_rulebook(idx.to(cuda_device), 1, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False)
    with pytest.raises(RuntimeError, match="duplicate"):
        ops.build_rulebook(idx.to(cuda_device), 1, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], True)
    # ... and through the layer API the reference calls
    import spconv
    layer = spconv.SparseConv2d(4, 4, 3, 1, 0, bias=False).to(cuda_device)
    x = spconv.SparseConvTensor(torch.ones(3, 4, device=cuda_device), idx.to(cuda_device), [14, 11], 1)
    with pytest.raises(RuntimeError, match="duplicate"):
        layer(x)


@pytest.mark.parametrize("k,s,p,d,subm", [(3, 1, 0, 1, False), (3, 1, 1, 1, False), (2, 1, 0, 1, False), (3, 2, 1, 1, False),
                                          (3, 1, 1, 1, True), (5, 1, 2, 1, True)])
def test_front_back_split_equals_whole(cuda_device, k, s, p, d, subm):
    """Graph path: the rulebook built in two launches (FRONT: output rows + nbr_out, what the forward pass waits for;
    BACK: pairs, nbr_in, counts, duplicate flag, beside the forward pass) is bit-identical to the single launch, on
    capacity-sized buffers with the live count on the device; duplicates are still flagged by the BACK half."""
    from waveformml_b200.spconv.functional import hints
    dev = cuda_device
    for dup in (False, True):
        idx = _indices(37, 11)
        n = idx.shape[0]
        if dup:
            idx[5] = idx[4]
        cap = n + 40
        buf = torch.zeros(cap, 3, dtype=torch.int32)
        buf[:n] = idx
        n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
        pad = [k // 2] * 2 if subm else [p, p]
        st = [1, 1] if subm else [s, s]
        res = []
        for front_only in (False, True):
            hints.start("record")
            try:
                rb = ops.build_rulebook(buf.to(dev), 37, [14, 11], [k, k], st, pad, [d, d], subm, n_rows=n_dev,
                                        front_only=front_only)
            finally:
                hints.stop()
            del ops.graph_dup_flags[:]
            if front_only:
                assert rb._pending, "the single-launch builder should have left its BACK half pending"
                front = (rb.outids.clone(), rb.nbr_out.clone(), rb.n_out_dev.clone())
                rb.finish()
                torch.cuda.synchronize()
                assert all(torch.equal(a, b) for a, b in zip(front, (rb.outids, rb.nbr_out, rb.n_out_dev))), \
                    "the BACK half must not touch the FRONT half's results"
            n_out = int(rb.n_out_dev)
            res.append((rb.outids[:n_out].cpu(), rb.nbr_out[:n_out].cpu(), rb.pairs[:, :, :n].cpu(), rb.pair_num.cpu(),
                        rb.nbr_in[:n].cpu(), int(rb.dup_flag.item()), n_out))
        whole, split = res
        if not dup:  # (with duplicate rows the contested nbr_out slots legitimately depend on which store lands last)
            for a, b in zip(whole, split):
                assert torch.equal(a, b) if torch.is_tensor(a) else a == b
        assert whole[5] == split[5] == (1 if dup else 0)
