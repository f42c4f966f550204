#!/usr/bin/env python
"""Times the 3x3 conv launches of the GEP stack alone (cold L2) at a given batch size, for every cluster split of the
(offset, slice) loop, with an exact-size grid and with a capacity-sized grid (graph path: live rows on the device).
Usage: exp_conv_small.py [BATCH]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waveformml_b200 import _lib
from waveformml_b200.spconv import ops
from waveformml_b200.spconv.fused import pitch8
from waveformml_b200.synth import make_events
dev = torch.device("cuda", 0)
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ev = make_events(B, n_samples=1, seed=1234)
idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(dev)
rb1 = ops.build_rulebook(idx, B, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False)
rb2 = ops.build_rulebook(rb1.outids, B, [12, 9], [3, 3], [1, 1], [0, 0], [1, 1], False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, nbr, n_src, n_dst, cin, cout, kvol, cap_factor):
    a16 = torch.randn(n_src, pitch8(cin), device=dev).to(torch.bfloat16)
    w = torch.randn(kvol, cin, cout, device=dev)
    wbuf = torch.empty(lib.wfsp_prepared_weight_bytes(kvol, cin, cout), dtype=torch.uint8, device=dev)
    job = (_lib.PrepJob * 1)(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr(), kvol, cin, cout, 0))
    _lib.check(lib.wfsp_prep_weights(ctypes.cast(job, ctypes.c_void_p), 1, _lib.stream()))
    cap = n_dst * cap_factor
    out = torch.empty(cap, cout, device=dev)
    nbr_cap = torch.full((cap, kvol), -1, dtype=torch.int32, device=dev)
    nbr_cap[:n_dst] = nbr
    n_dev = torch.tensor([n_dst], dtype=torch.int32, device=dev)
    res = []
    for ks in (1, 2, 4, 8):
        ep = _lib.conv_epilogue(k_split=ks)
        def run():
            _lib.check(lib.wfsp_conv_apply_bf16_ex(_lib.ptr(a16), n_src, None, cin, _lib.ptr(wbuf), None, _lib.ptr(nbr_cap), kvol,
                                                   _lib.ptr(out), cap, _lib.ptr(n_dev) if cap_factor > 1 else None,
                                                   n_dst if cap_factor > 1 else 0, cout, ctypes.byref(ep), _lib.stream()))
        ts = []
        run(); torch.cuda.synchronize()
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res.append("ks%d %.1f" % (ks, ts[len(ts) // 2]))
    print("%-22s rows %5d -> %5d, grid x%d: %s us (median, cold L2)" % (name, n_src, n_dst, cap_factor, "  ".join(res)))


n0, n1, n2 = idx.shape[0], rb1.outids.shape[0], rb2.outids.shape[0]
for cf in (1, 8):
    bench("L1 252->158 k3 fwd", rb1.nbr_out, n0, n1, 252, 158, 9, cf)
    bench("L2 158->64 k3 fwd", rb2.nbr_out, n1, n2, 158, 64, 9, cf)
    bench("L2 158->64 k3 dgrad", rb2.nbr_in, n2, n1, 64, 158, 9, cf)
    bench("L1 252->158 k3 dgrad", rb1.nbr_in, n1, n0, 158, 252, 9, cf)
