"""Window edges.  The oracle restatement is PINNED: tests/golden/window_edges_golden.npz
was produced by the reference's own C code (src/custom_functions/cffi.c compiled into oracle/_ref, see
tests/golden/make_window_edges_fixture.py).  CPU: oracle == golden.  GPU: kernel == golden == oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import spconv_cpu as osp

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "window_edges_golden.npz"))
CASES = sorted(k[:-6] for k in GOLD.files if k.endswith("_edges"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    dist, sl = (int(v) for v in GOLD[name + "_args"])
    got = osp.window_edges(GOLD[name + "_coo"], GOLD[name + "_batch"], dist, bool(sl))
    assert np.array_equal(got, GOLD[name + "_edges"])


def test_oracle_matches_compiled_reference_if_present():
    """In the build container the reference's C is compiled (oracle/_ref): compare on fresh random input."""
    import ctypes
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libcffi_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    ref = ctypes.CDLL(path)
    LLP = ctypes.POINTER(ctypes.c_longlong)
    rng = np.random.default_rng(0)
    n = 500
    coo = rng.integers(0, 14, size=(n, 2)).astype(np.int64)
    batch = np.sort(rng.integers(0, 60, size=n)).astype(np.int64)
    x, y = np.ascontiguousarray(coo[:, 0]), np.ascontiguousarray(coo[:, 1])
    e1, e2, cur = np.zeros(n * n, dtype=np.int64), np.zeros(n * n, dtype=np.int64), np.zeros(1, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(LLP)
    ref.cffi_window_edges(ctypes.c_longlong(3), p(cur), ctypes.c_int(n), p(x), p(y), p(batch), ctypes.c_bool(True), p(e1), p(e2))
    want = np.stack([e1[:cur[0]], e2[:cur[0]]], 0)
    assert np.array_equal(osp.window_edges(coo, batch, 2, True), want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_golden(cuda_device, name):
    from waveformml_b200.graph_utils import window_edges
    dist, sl = (int(v) for v in GOLD[name + "_args"])
    coo = torch.from_numpy(GOLD[name + "_coo"]).to(cuda_device)
    batch = torch.from_numpy(GOLD[name + "_batch"]).to(cuda_device)
    got = window_edges(coo, batch, max_dist=dist, self_loops=bool(sl))
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), GOLD[name + "_edges"])


@pytest.mark.gpu
def test_gpu_large_and_empty(cuda_device):
    """100 000 hits (several scan rounds) against the oracle; the empty input."""
    from waveformml_b200.graph_utils import window_edges
    rng = np.random.default_rng(1)
    n = 100000
    coo = rng.integers(0, 14, size=(n, 2)).astype(np.int64)
    batch = np.sort(rng.integers(0, 30000, size=n)).astype(np.int64)
    got = window_edges(torch.from_numpy(coo).to(cuda_device), torch.from_numpy(batch).to(cuda_device), 1, True)
    assert np.array_equal(got.cpu().numpy(), osp.window_edges(coo, batch, 1, True))
    e = window_edges(torch.zeros((0, 2), dtype=torch.int64, device=cuda_device),
                     torch.zeros((0,), dtype=torch.int64, device=cuda_device))
    assert tuple(e.shape) == (2, 0)
