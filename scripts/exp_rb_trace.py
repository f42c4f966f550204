#!/usr/bin/env python
"""Cycle trace of the single-launch rulebook builder (rb_small) on the two rulebooks of the 64-event GEP step,
graph-style arguments (capacity-sized buffers, live count on the device), cold L2 and warm L2."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waveformml_b200 import _lib
from waveformml_b200.spconv import ops
from waveformml_b200.spconv.functional import hints
from waveformml_b200.synth import make_events
dev = torch.device("cuda", 0)
lib = _lib.load()
B = 64
ev = make_events(B, n_samples=150, seed=1234)
idx = torch.from_numpy(ev["coords"])[:, [2, 0, 1]].contiguous().to(dev)
n = idx.shape[0]
cap = 640
idx_cap = torch.zeros(cap, 3, dtype=torch.int32, device=dev); idx_cap[:n] = idx
n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
trace = torch.zeros(128 + 2048, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["entry", "live count", "init", "claimed", "out rows", "nbr_out filled", "pairs", "exit"]
cur_idx, cur_n, shape = idx_cap, n_dev, [14, 11]
for layer in range(2):
    for cold in (True, False):
        for rep in range(3):
            if cold:
                flush.zero_()
            trace.zero_()
            hints.start("record")
            _lib.check(lib.wfsp_debug_trace(ctypes.c_void_p(trace.data_ptr())))
            rb = ops.build_rulebook(cur_idx, B, shape, [3, 3], [1, 1], [0, 0], [1, 1], False, n_rows=cur_n)
            torch.cuda.synchronize()
            _lib.check(lib.wfsp_debug_trace(None))
            hints.stop()
        t = trace.cpu()[:16]
        names2 = {8: "p2 taps", 9: "p2 scan sync", 10: "p2 assign", 11: "p3 count", 12: "p3 sync", 13: "p3 scan", 14: "p3 write"}
        print("   detail: " + "  ".join("%s %d" % (names2[i], int(t[i] - t[0])) for i in range(8, 15)))
        print("rulebook %d (%s L2), rows in %d -> out %d: " % (layer + 1, "cold" if cold else "warm", int(cur_n), int(rb.n_out_dev)) +
              "  ".join("%s %d" % (names[i], int(t[i] - t[0])) for i in range(1, 8)))
    cur_idx, cur_n, shape = rb.outids, rb.n_out_dev, [12, 9]
