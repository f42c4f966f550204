#!/usr/bin/env python
"""Times the 3x3 forward conv launch of the GEP stack alone (cold L2) at a given batch size."""
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
rb = ops.build_rulebook(idx, B, [14, 11], [3, 3], [1, 1], [0, 0], [1, 1], False)
n_in, n_out, cin, cout, kvol = idx.shape[0], rb.outids.shape[0], 252, 158, 9
a16 = torch.randn(n_in, pitch8(cin), device=dev).to(torch.bfloat16)
w = torch.randn(kvol, cin, cout, device=dev)
wbuf = torch.empty(lib.wfsp_prepared_weight_bytes(kvol, cin, cout), dtype=torch.uint8, device=dev)
job = (_lib.PrepJob * 1)(_lib.PrepJob(w.data_ptr(), wbuf.data_ptr(), kvol, cin, cout, 0))
_lib.check(lib.wfsp_prep_weights(ctypes.cast(job, ctypes.c_void_p), 1, _lib.stream()))
out = torch.empty(n_out, cout, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    _lib.check(lib.wfsp_conv_apply_bf16(_lib.ptr(a16), n_in, None, cin, _lib.ptr(wbuf), None, _lib.ptr(rb.nbr_out), kvol,
                                        _lib.ptr(out), n_out, None, 0, cout, None, _lib.stream()))
ts = []
run(); torch.cuda.synchronize()
for _ in range(20):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print("B %d rows in/out %d %d: 3x3 forward %.1f us (median, cold L2)" % (B, n_in, n_out, ts[len(ts) // 2]))
