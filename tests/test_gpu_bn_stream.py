"""BatchNorm(+ReLU) streaming passes fed by bulk copies (csrc/bn_stream.cu, large row counts) against the register-load
kernels of csrc/bn.cu (option "bn_stream" = 0) and against torch autograd: same arithmetic, different data path.
Row counts include an odd one whose last chunk is not a 16-byte multiple (copied by the threads) and a live count
below the capacity (graph path)."""
import pytest
import torch

from waveformml_b200 import _lib
from waveformml_b200.spconv.fused import pitch8

pytestmark = pytest.mark.gpu


def _fwd(lib, x, n, n_dev, c, gamma, beta, relu, part):
    dev = x.device
    y = torch.zeros(x.shape[0], c, device=dev)
    y16 = torch.zeros(x.shape[0], pitch8(c), dtype=torch.bfloat16, device=dev)
    mean, invstd = torch.empty(c, device=dev), torch.empty(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    _lib.check(lib.wfsp_bn_relu_fwd_stats(_lib.ptr(x), x.shape[0], _lib.ptr(n_dev), n, c, _lib.ptr(part), _lib.ptr(gamma),
                                          _lib.ptr(beta), _lib.ptr(rm), _lib.ptr(rv), 0.1, 1e-5, relu, _lib.ptr(y),
                                          _lib.ptr(y16), _lib.ptr(mean), _lib.ptr(invstd), _lib.stream()))
    return y, y16, mean, invstd


def _bwd(lib, x, dy, n, n_dev, c, gamma, beta, mean, invstd, relu):
    dev = x.device
    dx = torch.zeros(x.shape[0], c, device=dev)
    dx16 = torch.zeros(x.shape[0], pitch8(c), dtype=torch.bfloat16, device=dev)
    dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
    ws = torch.empty(lib.wfsp_bn_workspace_bytes(x.shape[0], c), dtype=torch.uint8, device=dev)
    _lib.check(lib.wfsp_bn_relu_bwd_x(_lib.ptr(x), _lib.ptr(dy), x.shape[0], _lib.ptr(n_dev), n, c, _lib.ptr(gamma),
                                      _lib.ptr(beta), _lib.ptr(mean), _lib.ptr(invstd), relu, _lib.ptr(dx), _lib.ptr(dx16),
                                      _lib.ptr(dg), _lib.ptr(db), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return dx, dx16, dg, db


@pytest.mark.parametrize("n,cap,c,relu", [(40001, 40001, 158, 1), (65536, 65536, 252, 1), (39000, 48000, 64, 0),
                                          (50003, 50003, 6, 1)])
def test_stream_equals_register_kernels(cuda_device, n, cap, c, relu):
    lib = _lib.load()
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(n + c)
    x = (torch.randn(cap, c, generator=g) * 1.5 + 0.3).to(dev)
    dy = torch.randn(cap, c, generator=g).to(dev)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(dev), (torch.randn(c, generator=g) * 0.2).to(dev)
    n_dev = torch.tensor([n], dtype=torch.int32, device=dev) if cap != n else None
    # per-32-row-chunk (mean, M2) partials as the convolution epilogue writes them
    chunks = (cap + 31) // 32
    part = torch.zeros(lib.wfsp_bn_partials_bytes(cap, c), dtype=torch.uint8, device=dev)
    pv = part[:chunks * 2 * c * 4].view(torch.float32).view(chunks, 2, c)
    xl = x[:n]
    for_chunks = (n + 31) // 32
    pad = for_chunks * 32 - n
    xp = torch.cat([xl, torch.zeros(pad, c, device=dev)]).view(for_chunks, 32, c)
    cnt = torch.full((for_chunks, 1), 32.0, device=dev)
    cnt[-1, 0] = 32 - pad
    mean_c = xp.sum(1) / cnt
    mask = (torch.arange(32, device=dev).view(1, 32, 1) < cnt.view(-1, 1, 1))
    pv[:for_chunks, 0] = mean_c
    pv[:for_chunks, 1] = (((xp - mean_c.unsqueeze(1)) ** 2) * mask).sum(1)
    res = {}
    for on in (1, 0):
        _lib.check(lib.wfsp_set_option(b"bn_stream", on))
        try:
            y, y16, mean, invstd = _fwd(lib, x, n, n_dev, c, gamma, beta, relu, part)
            res[on] = (y, y16, mean, invstd) + _bwd(lib, x, dy, n, n_dev, c, gamma, beta, mean, invstd, relu)
        finally:
            _lib.check(lib.wfsp_set_option(b"bn_stream", 1))
    for a, b, name in zip(res[1], res[0], ("y", "y16", "mean", "invstd", "dx", "dx16", "d_gamma", "d_beta")):
        a, b = a.float(), b.float()
        if a.dim() == 2:
            assert float(a[n:].abs().max()) == 0.0 if cap > n else True, name + ": rows beyond the live count written"
            a, b = a[:n], b[:n]
        # bf16 outputs: the two paths add the partial sums in different orders, which can move a value across a bf16
        # rounding boundary (one unit in the last place = 2^-8 relative)
        rtol = 8e-3 if name.endswith("16") else 2e-4
        torch.testing.assert_close(a, b, rtol=rtol, atol=2e-4 * max(float(b.abs().max()), 1e-6), msg=lambda m: name + ": " + m)
    # and against torch autograd (double)
    xr = xl.double().clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    yr = torch.nn.functional.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    if relu:
        yr = torch.relu(yr)
    yr.backward(dy[:n].double())
    torch.testing.assert_close(res[1][0][:n].double(), yr.detach(), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(res[1][4][:n].double(), xr.grad, rtol=2e-3, atol=2e-4 * float(xr.grad.abs().max()))
    torch.testing.assert_close(res[1][6].double(), gr.grad, rtol=2e-3, atol=2e-4 * float(gr.grad.abs().max()))
    torch.testing.assert_close(res[1][7].double(), br.grad, rtol=2e-3, atol=2e-4 * float(br.grad.abs().max()))
