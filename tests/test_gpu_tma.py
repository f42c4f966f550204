"""TMA gather4 (cp.async.bulk.tensor ... tile::gather4): four arbitrary rows of a bf16 matrix per instruction,
into 128-byte-swizzled shared memory.  Checks the tensor-map construction, the shared-memory layout the UMMA
producers assume, partial last channel slices, and that row indices outside the matrix read as zero (used
for missing neighbours)."""
import pytest
import torch

from waveformml_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,chan,c0", [(300, 64, 0), (1000, 200, 128), (77, 40, 0), (5000, 256, 192), (130, 8, 0)])
def test_gather4_matches_index_select(cuda_device, rows, chan, c0):
    lib = _lib.load()
    pitch = (chan + 7) // 8 * 8
    g = torch.Generator().manual_seed(rows)
    src = torch.zeros((rows, pitch), dtype=torch.bfloat16)
    src[:, :chan] = torch.randn((rows, chan), generator=g).to(torch.bfloat16)
    idx = torch.randint(0, rows, (128,), generator=g, dtype=torch.int32)
    idx[5], idx[6], idx[64], idx[127] = -1, rows, rows + 12345, 2 ** 31 - 1  # outside the matrix -> zeros
    d_src, d_idx = src.to(cuda_device), idx.to(cuda_device)
    out = torch.full((128, 64), 7.0, dtype=torch.bfloat16, device=cuda_device)
    with torch.cuda.device(cuda_device):
        _lib.check(lib.wfsp_selftest_gather4(_lib.ptr(d_src), rows, chan, pitch, _lib.ptr(d_idx), c0, _lib.ptr(out),
                                             _lib.stream()))
    torch.cuda.synchronize()
    ref = torch.zeros((128, 64), dtype=torch.bfloat16)
    for r in range(128):
        i = int(idx[r])
        if 0 <= i < rows:
            w = max(0, min(64, chan - c0))
            ref[r, :w] = src[i, c0:c0 + w]
    assert torch.equal(out.cpu(), ref)
