"""GPU batcher (wfsp_batch_pack) and dense scatter / gather (wfsp_to_dense[_bwd]) vs the oracle and
vs golden vectors produced by the reference's own collate_fn (tests/golden/batcher_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import spconv_cpu as osp
from waveformml_b200 import batcher, spconv
from waveformml_b200.synth import MAX_RANGE_INV, make_events

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "batcher_golden.npz")


def _items():
    g = np.load(GOLD)
    n = int(g["n_items"])
    coords = [g["coords%d" % i] for i in range(n)]
    wave = [g["wave%d" % i] for i in range(n)]
    labels = [g["labels%d" % i] for i in range(n)]
    return g, coords, wave, labels


def test_pack_matches_reference_collate_fn(cuda_device):
    g, coords, wave, labels = _items()
    rows = np.cumsum([0] + [c.shape[0] for c in coords]).tolist()
    evs = [l.shape[0] for l in labels]
    c = torch.from_numpy(np.concatenate(coords)).to(cuda_device)
    w = torch.from_numpy(np.concatenate(wave)).to(cuda_device)
    idx, feats = batcher.pack_batch(c, w, rows, evs)
    assert torch.equal(idx.cpu()[:, [1, 2, 0]], torch.from_numpy(g["out_coords"]))  # reference keeps (x, y, event)
    assert torch.equal(feats.cpu(), torch.from_numpy(g["out_feats"]))  # bit-exact fp32
    # bf16 output variant == rounding of the fp32 result
    _, fb = batcher.pack_batch(c, w, rows, evs, out_dtype=torch.bfloat16)
    assert torch.equal(fb.cpu(), torch.from_numpy(g["out_feats"]).bfloat16())
    # the collate_fn mirror (same call shape as the reference's)
    items = [([torch.from_numpy(ci).to(cuda_device), torch.from_numpy(wi).to(cuda_device)],
              torch.from_numpy(li).to(cuda_device)) for ci, wi, li in zip(coords, wave, labels)]
    (cc, ff), yy = batcher.collate_fn(items, scale=MAX_RANGE_INV)
    assert torch.equal(cc.cpu(), torch.from_numpy(g["out_coords"]))
    assert torch.equal(ff.cpu(), torch.from_numpy(g["out_feats"]))
    assert torch.equal(yy.cpu(), torch.from_numpy(g["out_labels"]))


@pytest.mark.parametrize("C", [300, 130, 7])
def test_pack_matches_oracle(cuda_device, C):
    ev = make_events(50, n_samples=1, seed=4)
    n = ev["coords"].shape[0]
    wave = torch.from_numpy(np.random.default_rng(1).integers(0, 2 ** 14, size=(n, C), dtype=np.int16))
    coords = torch.from_numpy(ev["coords"])
    oi, of, bs = osp.batch_pack(coords, wave, [0, n], [50], MAX_RANGE_INV)
    gi, gf = batcher.pack_batch(coords.to(cuda_device), wave.to(cuda_device))
    assert bs == 50 and torch.equal(gi.cpu(), oi) and torch.equal(gf.cpu(), of)


@pytest.mark.parametrize("shape,C", [((14, 11), 64), ((10, 7), 64), ((14, 11), 1), ((12, 9), 33)])
def test_to_dense_fwd_bwd(cuda_device, shape, C):
    B = 21
    ev = make_events(B, n_samples=1, seed=9)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()
    keep = (idx[:, 1] < shape[0]) & (idx[:, 2] < shape[1])
    idx = idx[keep].contiguous()
    feats = torch.randn(idx.shape[0], C)
    ref = osp.to_dense_c(feats, idx, B, shape)
    assert torch.equal(ref, osp.SparseConvTensor(feats, idx, shape, B).dense())
    fg = feats.to(cuda_device).requires_grad_(True)
    d = spconv.SparseConvTensor(fg, idx.to(cuda_device), list(shape), B).dense()
    assert torch.equal(d.cpu(), ref)
    g = torch.randn(ref.shape)
    (d * g.to(cuda_device)).sum().backward()
    li = idx.long()
    assert torch.equal(fg.grad.cpu(), g[li[:, 0], :, li[:, 1], li[:, 2]])
    nhwc = spconv.SparseConvTensor(fg.detach(), idx.to(cuda_device), list(shape), B).dense(channels_first=False)
    assert torch.equal(nhwc.cpu(), ref.permute(0, 2, 3, 1))


def test_to_dense_two_halves_equal_whole(cuda_device):
    """wfsp_dense_cell_table (on another stream, as the fused stack builds it beside the convolutions) +
    wfsp_to_dense_from_table == wfsp_to_dense, bit for bit, with duplicate coordinates (last row wins)."""
    from waveformml_b200 import _lib
    lib = _lib.load()
    B, shape, C = 9, (14, 11), 45
    ev = make_events(B, n_samples=1, seed=3)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous()
    idx = torch.cat([idx, idx[:5]]).contiguous()  # five duplicates
    n = idx.shape[0]
    feats = torch.randn(n, C)
    ref = osp.to_dense_c(feats, idx, B, shape)
    fg, ig = feats.to(cuda_device), idx.to(cuda_device)
    whole = torch.empty((B, C) + shape, device=cuda_device)
    halves = torch.empty_like(whole)
    t1 = torch.empty((B * shape[0] * shape[1],), dtype=torch.int32, device=cuda_device)
    t2 = torch.empty_like(t1)
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    torch.cuda.synchronize()
    _lib.check(lib.wfsp_to_dense(_lib.ptr(fg), _lib.ptr(ig), n, _lib.ptr(n_dev), C, B, shape[0], shape[1], _lib.ptr(whole),
                                 _lib.ptr(t1), _lib.stream()))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        _lib.check(lib.wfsp_dense_cell_table(_lib.ptr(ig), n, _lib.ptr(n_dev), B, shape[0], shape[1], _lib.ptr(t2),
                                             _lib.stream()))
    torch.cuda.current_stream().wait_stream(side)
    _lib.check(lib.wfsp_to_dense_from_table(_lib.ptr(fg), C, B, shape[0], shape[1], _lib.ptr(t2), _lib.ptr(halves),
                                            _lib.stream()))
    assert torch.equal(t1.cpu(), t2.cpu())
    assert torch.equal(whole.cpu(), ref) and torch.equal(halves.cpu(), ref)


def test_to_dense_empty(cuda_device):
    d = spconv.SparseConvTensor(torch.zeros(0, 3, device=cuda_device), torch.zeros(0, 3, dtype=torch.int32, device=cuda_device),
                                [14, 11], 2).dense()
    assert d.shape == (2, 3, 14, 11) and float(d.abs().sum()) == 0.0


def test_pack_bf16_operand_format_padding(cuda_device):
    """bf16 output = tensor-core operand format: pitch rounded up to 8 channels, padding zero (both kernels:
    the 4-wide vector path c % 4 == 0 and the scalar path)."""
    for c in (12, 300, 7):
        g = torch.Generator().manual_seed(c)
        n = 37
        wave = torch.randint(0, 2 ** 14, (n, c), generator=g, dtype=torch.int16).to(cuda_device)
        coords = torch.zeros((n, 3), dtype=torch.int32, device=cuda_device)
        _, f32 = batcher.pack_batch(coords, wave, [0, n], [1])
        _, f16 = batcher.pack_batch(coords, wave, [0, n], [1], out_dtype=torch.bfloat16)
        pitch = (c + 7) // 8 * 8
        assert tuple(f16.shape) == (n, pitch)
        assert torch.equal(f16[:, :c], f32.bfloat16())
        assert bool((f16[:, c:] == 0).all())


def test_stage_inputs_copies_bit_exact(cuda_device):
    """wfsp_stage_inputs: three copies + the live row count in one launch; odd byte counts, unaligned
    pointers (a slice starting at an odd element), an empty and a skipped slot."""
    from waveformml_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    src = [torch.randint(-2 ** 15, 2 ** 15, (4097,), dtype=torch.int16, generator=g).to(cuda_device)[1:],  # 2-byte aligned only
           torch.randint(0, 255, (100003,), dtype=torch.uint8, generator=g).to(cuda_device),
           torch.randint(0, 2 ** 31 - 1, (777, 3), dtype=torch.int32, generator=g).to(cuda_device)]
    dst = [torch.full((s.numel() + 9,), 7, dtype=s.dtype, device=cuda_device) for s in src]
    n_dev = torch.zeros((1,), dtype=torch.int32, device=cuda_device)
    with torch.cuda.device(cuda_device):
        _lib.check(lib.wfsp_stage_inputs(_lib.ptr(dst[0]), _lib.ptr(src[0]), src[0].numel() * 2,
                                         _lib.ptr(dst[1]), _lib.ptr(src[1]), src[1].numel(),
                                         _lib.ptr(dst[2]), _lib.ptr(src[2]), src[2].numel() * 4,
                                         _lib.ptr(n_dev), 4321, _lib.stream()))
    for d, s in zip(dst, src):
        assert torch.equal(d[:s.numel()], s.reshape(-1)) and bool((d[s.numel():] == 7).all())
    assert int(n_dev.item()) == 4321
    with torch.cuda.device(cuda_device):  # skipped slots, count only
        _lib.check(lib.wfsp_stage_inputs(None, None, 0, _lib.ptr(dst[1]), None, 5, None, None, 0, _lib.ptr(n_dev), 5,
                                         _lib.stream()))
    assert int(n_dev.item()) == 5 and torch.equal(dst[1][:src[1].numel()], src[1])
