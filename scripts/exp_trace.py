#!/usr/bin/env python
"""Cycle trace of tile (0,0) of the L1 3x3 forward launch at 64 events, per cluster split (wfsp_debug_trace)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waveformml_b200 import _lib
from waveformml_b200.spconv import ops
from waveformml_b200.spconv.fused import pitch8
from waveformml_b200.synth import make_events
dev = torch.device("cuda", 0)
lib = _lib.load()
B = 64
ev = make_events(B, n_samples=1, seed=1234)
idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(dev)
rb1 = ops.build_rulebook(idx, B, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False)
n_src, n_dst, cin, cout, kvol = idx.shape[0], rb1.outids.shape[0], 252, 158, 9
a16 = torch.randn(n_src, pitch8(cin), device=dev).to(torch.bfloat16)
w = torch.randn(kvol, cin, cout, device=dev)
wbuf = torch.empty(lib.wfsp_prepared_weight_bytes(kvol, cin, cout), dtype=torch.uint8, device=dev)
job = (_lib.PrepJob * 1)(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr(), kvol, cin, cout, 0))
_lib.check(lib.wfsp_prep_weights(ctypes.cast(job, ctypes.c_void_p), 1, _lib.stream()))
out = torch.empty(n_dst, cout, device=dev)
trace = torch.zeros(128 + 2048, dtype=torch.int64, device=dev)
names = ["entry", "prologue", "nbr tile", "producers done", "mma done", "parked", "cluster sync 1", "finished blocks", "cluster sync 2", "exit"]
for kv in os.environ.get("WFSP_OPTIONS", "").split(","):
    if kv:
        k, v = kv.split("=")
        _lib.check(lib.wfsp_set_option(k.encode(), int(v)))
for ks in (1, 2, 4, 8):
    ep = _lib.conv_epilogue(k_split=ks)
    for warm in range(3):
        trace.zero_()
        _lib.check(lib.wfsp_debug_trace(ctypes.c_void_p(trace.data_ptr())))
        _lib.check(lib.wfsp_conv_apply_bf16_ex(_lib.ptr(a16), n_src, None, cin, _lib.ptr(wbuf), None, _lib.ptr(rb1.nbr_out), kvol,
                                               _lib.ptr(out), n_dst, None, 0, cout, ctypes.byref(ep), _lib.stream()))
        torch.cuda.synchronize()
    _lib.check(lib.wfsp_debug_trace(None))
    tc = trace.cpu()
    t = tc[:128].view(8, 16)
    g0 = int(t[:ks, 15].min())
    se = tc[128:].view(-1, 2)
    live = se[:, 0] > 0
    if bool(live.any()):
        st, en = se[live, 0], se[live, 1]
        print("ks %d: %d live CTAs; entry spread %d ns, first entry -> last exit %d ns, mean CTA life %d ns" % (
            ks, int(live.sum()), int(st.max() - st.min()), int(en.max() - st.min()), int((en - st).float().mean())))
    print("ks %d (warm L2): per rank, cycles since the CTA's entry; entry offset in ns from the first rank" % ks)
    for r in range(ks):
        row = t[r]
        print("  rank %d entry +%5d ns: " % (r, int(row[15]) - g0) + "  ".join(
            "%s %d" % (names[i], int(row[i] - row[0])) for i in range(1, 10) if int(row[i]) != 0))
