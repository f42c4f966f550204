"""Extended epilogues / launch options of the tcgen05 apply kernel, through the C ABI (wfsp_conv_apply_bf16_ex):

  * k_split: the (kernel offset x channel slice) loop of a tile split over a thread-block cluster, partial accumulators
    added through distributed shared memory in rank order -- must equal the unsplit launch up to fp32 re-association,
    give the same BatchNorm partial statistics, and be run-to-run bit-identical (fixed summation order);
  * bwd_partials: the two reductions of BatchNorm(+ReLU) backward taken from the dgrad tile, and
    wfsp_bn_relu_bwd_parts against the two-pass wfsp_bn_relu_bwd_x and against torch autograd."""
import ctypes

import pytest
import torch

from waveformml_b200 import _lib
from waveformml_b200.spconv import ops
from waveformml_b200.spconv.fused import pitch8
from waveformml_b200.synth import make_events

pytestmark = pytest.mark.gpu


def _setup(dev, B, cin, cout, k, full=False, seed=5):
    lib = _lib.load()
    ev = make_events(B, n_samples=1, seed=seed, full_grid=full)
    idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(dev)
    if k > 1:
        rb = ops.build_rulebook(idx, B, [14, 11], [k, k], [1, 1], [0, 0], [1, 1], False)
        nbr, n_out, kvol = rb.nbr_out, rb.outids.shape[0], k * k
    else:
        nbr, n_out, kvol = None, idx.shape[0], 1
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_in = idx.shape[0]
    a = torch.randn(n_in, cin, generator=g).to(dev)
    a16 = torch.zeros(n_in, pitch8(cin), dtype=torch.bfloat16, device=dev)
    a16[:, :cin] = a.to(torch.bfloat16)
    w = (torch.randn(kvol, cin, cout, generator=g) / (cin * kvol) ** 0.5).to(dev)
    wbuf = torch.empty(lib.wfsp_prepared_weight_bytes(kvol, cin, cout), dtype=torch.uint8, device=dev)
    job = (_lib.PrepJob * 1)(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr(), kvol, cin, cout, 0))
    _lib.check(lib.wfsp_prep_weights(ctypes.cast(job, ctypes.c_void_p), 1, _lib.stream()))
    return lib, a16, w, wbuf, nbr, n_in, n_out, kvol


def _apply(lib, a16, wbuf, nbr, n_in, n_out, kvol, cin, cout, ep, bias=None):
    out = torch.full((n_out, cout), float("nan"), device=a16.device)
    _lib.check(lib.wfsp_conv_apply_bf16_ex(_lib.ptr(a16), n_in, None, cin, _lib.ptr(wbuf), _lib.ptr(bias), _lib.ptr(nbr),
                                           kvol, _lib.ptr(out), n_out, None, 0, cout, ctypes.byref(ep), _lib.stream()))
    return out


@pytest.mark.parametrize("B,cin,cout,k", [(64, 252, 158, 3), (64, 158, 64, 3), (16, 300, 252, 1), (40, 20, 300, 3),
                                          (7, 64, 5, 3)])
def test_k_split_matches_unsplit(cuda_device, B, cin, cout, k):
    lib, a16, w, wbuf, nbr, n_in, n_out, kvol = _setup(cuda_device, B, cin, cout, k)
    bias = torch.randn(cout, device=cuda_device)
    outs, stats = {}, {}
    for ks in (1, 2, 4, 8):
        part = torch.zeros(lib.wfsp_bn_partials_bytes(n_out, cout), dtype=torch.uint8, device=cuda_device)
        ep = _lib.conv_epilogue(bn_partials=part, k_split=ks)
        outs[ks] = _apply(lib, a16, wbuf, nbr, n_in, n_out, kvol, cin, cout, ep, bias)
        chunks = (n_out + 31) // 32
        stats[ks] = part[:chunks * 2 * cout * 4].view(torch.float32).clone()
        again = _apply(lib, a16, wbuf, nbr, n_in, n_out, kvol, cin, cout, ep, bias)
        assert torch.equal(outs[ks], again), "k_split %d is not run-to-run deterministic" % ks
    assert not torch.isnan(outs[1]).any()
    scale = float(outs[1].abs().max())
    for ks in (2, 4, 8):
        torch.testing.assert_close(outs[ks], outs[1], rtol=1e-5, atol=1e-5 * scale)
        torch.testing.assert_close(stats[ks], stats[1], rtol=1e-3, atol=1e-4 * scale * scale)


def test_auto_split_equals_forced_off(cuda_device):
    """The automatic choice (small launch -> cluster split) against the option that switches it off."""
    lib, a16, w, wbuf, nbr, n_in, n_out, kvol = _setup(cuda_device, 64, 252, 158, 3)
    ep = _lib.conv_epilogue()
    auto = _apply(lib, a16, wbuf, nbr, n_in, n_out, kvol, 252, 158, ep)
    _lib.check(lib.wfsp_set_option(b"apply_k_split", 0))
    try:
        off = _apply(lib, a16, wbuf, nbr, n_in, n_out, kvol, 252, 158, ep)
    finally:
        _lib.check(lib.wfsp_set_option(b"apply_k_split", 1))
    torch.testing.assert_close(auto, off, rtol=1e-5, atol=1e-5 * float(off.abs().max()))


@pytest.mark.parametrize("B,full,c,relu,ks", [(64, False, 158, 1, 0), (64, False, 252, 0, 1), (96, True, 64, 1, 0),
                                              (5, False, 3, 1, 4)])
def test_bn_backward_from_dgrad_partials(cuda_device, B, full, c, relu, ks):
    """dgrad epilogue partial sums + wfsp_bn_relu_bwd_parts == two-pass wfsp_bn_relu_bwd_x == torch autograd."""
    dev = cuda_device
    cred = 40
    # a '1x1 dgrad': dy = g16 @ W (identity rulebook), arriving at BatchNorm(c)+ReLU whose input was x
    lib, g16, w, wbuf, nbr, n, _, kvol = _setup(dev, B, cred, c, 1, full=full, seed=9)
    gen = torch.Generator(device="cpu").manual_seed(3)
    x = (torch.randn(n, c, generator=gen) * 2 + 0.5).to(dev)
    gamma, beta = (torch.rand(c, generator=gen) + 0.5).to(dev), (torch.randn(c, generator=gen) * 0.3).to(dev)
    mean = x.mean(0)
    invstd = 1.0 / torch.sqrt(x.var(0, unbiased=False) + 1e-5)
    parts = torch.zeros(lib.wfsp_bn_partials_bytes(n, c), dtype=torch.uint8, device=dev)
    ep = _lib.conv_epilogue(bwd=(x, mean, invstd, gamma, beta, relu, parts), k_split=ks)
    dy = _apply(lib, g16, wbuf, None, n, n, 1, cred, c, ep)
    # partial sums against a direct evaluation
    xh = (x - mean) * invstd
    dym = dy * ((xh * gamma + beta) > 0) if relu else dy
    chunks = (n + 31) // 32
    p = parts[:chunks * 2 * c * 4].view(torch.float32).view(chunks, 2, c)
    pad = chunks * 32 - n
    dpad = torch.cat([dym, torch.zeros(pad, c, device=dev)]).view(chunks, 32, c)
    xpad = torch.cat([xh, torch.zeros(pad, c, device=dev)]).view(chunks, 32, c)
    torch.testing.assert_close(p[:, 0], dpad.sum(1), rtol=1e-4, atol=1e-4 * float(dy.abs().max()))
    torch.testing.assert_close(p[:, 1], (dpad * xpad).sum(1), rtol=1e-4, atol=1e-4 * float((dy * xh).abs().max()))

    def run(parts_path):
        dx = torch.empty(n, c, device=dev)
        dx16 = torch.empty(n, pitch8(c), dtype=torch.bfloat16, device=dev)
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
        if parts_path:
            _lib.check(lib.wfsp_bn_relu_bwd_parts(_lib.ptr(x), _lib.ptr(dy), n, None, 0, c, _lib.ptr(gamma), _lib.ptr(beta),
                                                  _lib.ptr(mean), _lib.ptr(invstd), relu, _lib.ptr(parts), _lib.ptr(dx),
                                                  _lib.ptr(dx16), _lib.ptr(dg), _lib.ptr(db), _lib.stream()))
        else:
            ws = torch.empty(lib.wfsp_bn_workspace_bytes(n, c), dtype=torch.uint8, device=dev)
            _lib.check(lib.wfsp_bn_relu_bwd_x(_lib.ptr(x), _lib.ptr(dy), n, None, 0, c, _lib.ptr(gamma), _lib.ptr(beta),
                                              _lib.ptr(mean), _lib.ptr(invstd), relu, _lib.ptr(dx), _lib.ptr(dx16),
                                              _lib.ptr(dg), _lib.ptr(db), _lib.ptr(ws), ws.numel(), _lib.stream()))
        return dx, dx16, dg, db

    a, b = run(True), run(False)
    for u, v in zip(a, b):
        torch.testing.assert_close(u.float(), v.float(), rtol=2e-4, atol=2e-4 * max(float(v.float().abs().max()), 1e-6))
    # torch autograd of the same BatchNorm(+ReLU) in double
    xr = x.double().clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    y = torch.nn.functional.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    if relu:
        y = torch.relu(y)
    y.backward(dy.double())
    torch.testing.assert_close(a[0].double(), xr.grad, rtol=1e-3, atol=1e-4 * float(xr.grad.abs().max()))
    torch.testing.assert_close(a[2].double(), gr.grad, rtol=1e-3, atol=1e-4 * float(gr.grad.abs().max()))
    torch.testing.assert_close(a[3].double(), br.grad, rtol=1e-3, atol=1e-4 * float(br.grad.abs().max()))
