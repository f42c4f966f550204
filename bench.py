#!/usr/bin/env python
"""bench.py -- sparse-conv forward+backward events/s of the PSD classifier (GEP.json stack) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload W]

One "step" = one training step over one synthetic batch of B events per GPU: batcher (int16 -> f32,
event offsets, batch-first permute) -> rulebooks -> 3 sparse convs + BN/ReLU -> ToDense -> Linear x2 ->
CrossEntropy -> backward (dgrad + wgrad) -> flat-gradient all-reduce (N > 1) -> SGD update.

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MB write) before each
timed step, barrier + synchronize on both sides of the timed loop, max over ranks.  `value` has the
batch resident in HBM; `e2e` starts from pinned host buffers (H2D inside the timed region) and ends
with the loss read back to the host.

Besides the contract keys the line carries: `roofline` (dominant conv launch of the headline workload, timed
alone, against the BURST tensor peak / the copy bandwidth), `large` (the throughput-bound configuration C5 =
1024 full-grid events per GPU: ms_per_step, value, e2e, its own roofline and per-kernel fractions),
`sustained` (>= 2000 replays), `rotating` (8 distinct batches with different row counts through the one
captured graph), `cpu_baseline`, `clocks`, `gpu_launches`.

--impl reference times the reference's CPU path (restated spconv-1.2.1 CPU algorithm: oracle/) on the
host cores, same model / batch / metric.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "sparse-conv fwd+bwd events/sec"
WORKLOADS = {
    # name: (description, full_grid)
    "C2": ("C2: LitPSD PSD classifier (GEP.json stack 300->252 k1, 252->158 k3, 158->64 k3, ToDense, "
           "Linear 4480->116->3), synthetic 14x11 events, 2x150 samples/hit, clustered 1-10 hits/event", False),
    "C5": ("C5: same PSD classifier on high-occupancy events (all 154 cells hit)", True),
}


_RESULT = None  # the process's real stdout (see _quiet_stdout)


def emit(line):
    out = _RESULT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64, help="events per GPU")
    ap.add_argument("--math", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--no-math-modes", action="store_true", help="skip the bf16x3 / fp32 figures of the default line")
    ap.add_argument("--c3-batch", type=int, default=1024, help="events of the C3 (z-regression) block, 0 = skip")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: sync-free capacity-sized step replayed from a CUDA graph; eager: exact shapes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--large-batch", type=int, default=1024,
                    help="also time the throughput-bound configuration (C5, this many full-grid events per GPU) and "
                         "report it under `large`; 0 = skip")
    ap.add_argument("--sustained", type=int, default=2000, help="replays of the `sustained` figure; 0 = skip")
    ap.add_argument("--rotate", type=int, default=8, help="distinct batches of the `rotating` figure; <= 1 = skip")
    return ap.parse_args()


def config_of(args, world, rows):
    """`config` names the WORKLOAD only and is identical in both arms (ours / reference); how each arm executes it
    is under `implementation` / `cpu_baseline`."""
    return {"workload": WORKLOADS[args.workload][0], "events_per_gpu": args.batch,
            "global_events_per_step": world * args.batch, "rows_per_gpu": int(rows),
            "l2": "GPU arm: flushed with a 256 MB write before every timed step"}


def make_batch(args, rank):
    from waveformml_b200.synth import make_events
    full = WORKLOADS[args.workload][1]
    return make_events(args.batch, n_samples=150, seed=1234 + rank, full_grid=full)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
def cpu_reference_model(model):
    """Oracle twin (restated spconv-1.2.1 CPU algorithm) of the GPU model, weights shared."""
    import copy
    from oracle import mirror
    sparse = mirror.to_oracle(model.sparseModel).train()
    linear = copy.deepcopy(model.linear).cpu().train()
    return sparse, linear


def cpu_step_fn(model_cpu_state, batch, n_linear):
    from oracle import mirror
    from oracle import spconv_cpu as osp
    from waveformml_b200.synth import MAX_RANGE_INV
    sparse, linear = model_cpu_state
    params = list(sparse.parameters()) + list(linear.parameters())
    opt = torch.optim.SGD(params, lr=0.02, momentum=0.98, nesterov=True)
    crit = torch.nn.CrossEntropyLoss()
    coords, wave = torch.from_numpy(batch["coords"]), torch.from_numpy(batch["wave"])
    labels = torch.from_numpy(batch["labels"])
    n_ev = labels.shape[0]

    def step():
        idx, feats, bs = osp.batch_pack(coords, wave, [0, coords.shape[0]], [n_ev], MAX_RANGE_INV)
        opt.zero_grad(set_to_none=True)
        d = mirror.run_stack(sparse, idx, feats, [14, 11], bs)
        loss = crit(linear(d.view(-1, n_linear)), labels)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def pick_cpu_threads(step):
    """The CPU path is many small GEMMs and gathers: more threads is not faster (on a 100+-core host all cores
    is several times SLOWER than a handful).  To time the reference at its best, try a few thread counts on
    one step each and keep the fastest; the count used is reported as `cores`."""
    ncpu = os.cpu_count() or 1
    best_t, best_s = 1, None
    for t in sorted({1, 4, 16, min(64, ncpu), ncpu}):
        if t > ncpu:
            continue
        torch.set_num_threads(t)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if best_s is None or dt < best_s:
            best_t, best_s = t, dt
    torch.set_num_threads(best_t)
    return best_t


def time_cpu(step, budget_s, warmup=2, min_steps=3, max_steps=200):
    for _ in range(warmup):
        step()
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_steps or (time.perf_counter() < t_end and len(times) < max_steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from waveformml_b200 import stacks
    torch.manual_seed(0)
    model = stacks.PSDClassifier()
    batch = make_batch(args, 0)
    step = cpu_step_fn(cpu_reference_model(model), batch, model.n_linear)
    pick_cpu_threads(step)
    for _ in range(max(args.warmup, 1)):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if sum(times) > 240:  # keep the whole run within a few minutes
            break
    ms = 1e3 * float(np.mean(times))
    value = args.batch / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, int(os.environ.get("WORLD_SIZE", "1")), batch["coords"].shape[0]),
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d full training steps of ONE rank's %d events on one process (restated spconv-1.2.1 "
                                   "CPU algorithm: C hash rulebook + per-offset gather/torch.mm/scatter-add, torch-CPU "
                                   "BN/ReLU/Linear/SGD); thread count = fastest of {1,4,16,64,all} on this host (%d cores)"
                                   % (len(times), args.batch, os.cpu_count() or 1)},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def conv_breakdown(model, idx, feats, batch_size, flush, reps=10):
    """Times every sparse-conv kernel call (forward, dgrad, wgrad) of the model separately with CUDA
    events on the launching stream and returns per-call algorithmic FLOPs / bytes."""
    from waveformml_b200 import spconv
    from waveformml_b200.spconv import functional as Fsp
    captured = []

    def hook(mod, inp, out):
        x = inp[0]
        rb = None if mod.conv1x1 else x.indice_dict[mod.indice_key]
        captured.append((mod, x.features.detach(), rb, out.features.detach()))

    hooks = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, spconv.SparseConvolution)]
    with torch.no_grad():
        model([idx, feats, batch_size])
    for h in hooks:
        h.remove()
    mode = spconv.get_math_mode()
    e = 2 if mode == "bf16" else 4

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts)) * 1e-3

    import ctypes
    from waveformml_b200 import _lib
    from waveformml_b200.spconv.fused import pitch8
    lib = _lib.load()
    st = _lib.stream

    def cast16(t):
        out = torch.empty((t.shape[0], pitch8(t.shape[1])), dtype=torch.bfloat16, device=t.device)
        _lib.check(lib.wfsp_cast_rows_bf16(_lib.ptr(t), t.shape[0], None, t.shape[1], _lib.ptr(out), st()))
        return out

    def prep(w3, kvol, c_red, c_dst, transpose):
        buf = torch.empty((lib.wfsp_prepared_weight_bytes(kvol, c_red, c_dst),), dtype=torch.uint8, device=w3.device)
        job = (_lib.PrepJob * 1)(_lib.PrepJob(w3.data_ptr(), buf.data_ptr(), kvol, c_red, c_dst, transpose))
        _lib.check(lib.wfsp_prep_weights(ctypes.cast(job, ctypes.c_void_p), 1, st()))
        return buf

    rows = []
    for li, (mod, fin, rb, fout) in enumerate(captured):
        kvol = 1 if rb is None else rb.kvol
        cin, cout = mod.in_channels, mod.out_channels
        w3 = mod.weight.detach().view(kvol, cin, cout).contiguous()
        n_in, n_out = fin.shape[0], fout.shape[0]
        pairs = n_in if rb is None else int(rb.pair_num.sum().item())
        g = torch.randn_like(fout)
        a16, g16 = cast16(fin.contiguous()), cast16(g)
        w_f, w_d = prep(w3, kvol, cin, cout, 0), prep(w3, kvol, cout, cin, 1)
        out_f = torch.empty((n_out, cout), device=fin.device)
        out_d = torch.empty((n_in, cin), device=fin.device)
        dw = torch.empty((kvol, cin, cout), device=fin.device)
        flops = 2.0 * pairs * cin * cout
        # algorithmic traffic of ONE kernel launch: bf16 operand rows read once, fp32
        # result rows written once, bf16 weights read once (wgrad: fp32 d_weight written once), 4 B of
        # neighbour index per pair (wgrad: 8 B, both sides of the pair)
        fwd_b = 2 * n_in * pitch8(cin) + 4 * n_out * cout + 2 * kvol * cin * cout + 4 * pairs
        dg_b = 2 * n_out * pitch8(cout) + 4 * n_in * cin + 2 * kvol * cin * cout + 4 * pairs
        wg_b = 2 * n_in * pitch8(cin) + 2 * n_out * pitch8(cout) + 4 * kvol * cin * cout + 8 * pairs
        nbr_o = None if rb is None else rb.nbr_out
        nbr_i = None if rb is None else rb.nbr_in
        pa = None if rb is None else rb.pairs[0]
        pb = None if rb is None else rb.pairs[1]
        pn = None if rb is None else rb.pair_num
        pitch = n_in if rb is None else rb.pairs.shape[-1]
        name = "L%d %d->%d k%d" % (li, cin, cout, mod.kernel_size[0])

        def fwd():
            _lib.check(lib.wfsp_conv_apply_bf16(_lib.ptr(a16), n_in, None, cin, _lib.ptr(w_f), None, _lib.ptr(nbr_o), kvol,
                                                _lib.ptr(out_f), n_out, None, 0, cout, None, st()))

        def dgrad():
            _lib.check(lib.wfsp_conv_apply_bf16(_lib.ptr(g16), n_out, None, cout, _lib.ptr(w_d), None, _lib.ptr(nbr_i), kvol,
                                                _lib.ptr(out_d), n_in, None, 0, cin, None, st()))

        def wgrad():
            _lib.check(lib.wfsp_conv_wgrad_bf16(_lib.ptr(a16), n_in, None, cin, _lib.ptr(g16), n_out, None, cout,
                                                _lib.ptr(pa), _lib.ptr(pb), _lib.ptr(pn), kvol, pitch, 0, _lib.ptr(dw), 0,
                                                st()))

        rows.append({"name": name + " fwd", "kernel": "conv_apply_umma_kernel", "s": timed(fwd), "flops": flops, "bytes": fwd_b})
        if li > 0:
            rows.append({"name": name + " dgrad", "kernel": "conv_apply_umma_kernel", "s": timed(dgrad), "flops": flops, "bytes": dg_b})
        rows.append({"name": name + " wgrad", "kernel": "conv_wgrad_umma_kernel", "s": timed(wgrad), "flops": flops, "bytes": wg_b})
    return rows, e


def bn_breakdown(rows_c, flush, pk, reps=10):
    """The HBM-bound passes of the step (BatchNorm + ReLU between the convolutions, SPConvBlocks.py:505-508) timed alone
    with CUDA events, L2 flushed: achieved GB/s over ALGORITHMIC bytes (forward: 4 B read + 2 B bf16 written per
    element; backward: x and dy read twice (reduction, then update) + 2 B written = 18 B per element)."""
    from waveformml_b200 import _lib
    from waveformml_b200.spconv.fused import pitch8
    lib = _lib.load()
    out = []
    for name, n, c in rows_c:
        dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.randn(n, c, device=dev)
        dy = torch.randn(n, c, device=dev)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        y16 = torch.empty(n, pitch8(c), dtype=torch.bfloat16, device=dev)
        part = torch.zeros(lib.wfsp_bn_partials_bytes(n, c), dtype=torch.uint8, device=dev)
        ws = torch.empty(lib.wfsp_bn_workspace_bytes(n, c), dtype=torch.uint8, device=dev)
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)

        def fwd():
            _lib.check(lib.wfsp_bn_relu_fwd_stats(_lib.ptr(x), n, None, 0, c, _lib.ptr(part), _lib.ptr(gamma), _lib.ptr(beta),
                                                  _lib.ptr(rm), _lib.ptr(rv), 0.1, 1e-5, 1, None, _lib.ptr(y16),
                                                  _lib.ptr(mean), _lib.ptr(invstd), _lib.stream()))

        def bwd():
            _lib.check(lib.wfsp_bn_relu_bwd_x(_lib.ptr(x), _lib.ptr(dy), n, None, 0, c, _lib.ptr(gamma), _lib.ptr(beta),
                                              _lib.ptr(mean), _lib.ptr(invstd), 1, None, _lib.ptr(y16), _lib.ptr(dg),
                                              _lib.ptr(db), _lib.ptr(ws), ws.numel(), _lib.stream()))

        for what, fn, bpe in (("BatchNorm+ReLU forward (fold + normalise)", fwd, 6), ("BatchNorm+ReLU backward (reduce + fold + update)", bwd, 18)):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            sec = float(np.mean(ts))
            by = float(n) * c * bpe
            out.append({"name": "%s %s [%d x %d]" % (name, what, n, c), "ms": sec * 1e3, "algorithmic_bytes": by,
                        "gbs": by / sec / 1e9, "frac_hbm": by / sec / 1e9 / pk["hbm_gbs"]})
    return out


def roofline_of(row, pk, ms_per_step, traffic_key, note=None):
    """roofline object for ONE isolated kernel launch (CUDA events, L2 flushed): the denominator is the BURST
    tensor peak (a kernel timed alone), `frac_sustained` rides along; HBM-bound launches use the copy bandwidth."""
    ai = row["flops"] / row["bytes"]
    ridge = pk["tflops_burst"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    tf, gb = row["flops"] / row["s"] / 1e12, row["bytes"] / row["s"] / 1e9
    if ai >= ridge:
        roof = {"bound": "tensor", "achieved": tf, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                "frac": tf / pk["tflops_burst"], "frac_sustained": tf / pk["tflops_sustained"]}
    else:
        roof = {"bound": "hbm", "achieved": gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / pk["hbm_gbs"]}
    # DRAM bytes of the same launch from the committed `ncu --set full` capture (profiles/r*_traffic.json, made by
    # scripts/ncu_traffic.py); null if that capture does not exist
    roof["traffic"] = None
    roof["algorithmic_bytes"] = row["bytes"]
    roof["algorithmic_flops"] = row["flops"]
    for fname in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", fname)))
            parts = row["name"].split()
            ent = tr[traffic_key][parts[0] + " " + parts[-1]]
            roof["traffic"] = ent["dram_bytes"]
            roof["traffic_source"] = "profiles/%s (ncu --set full, %s)" % (fname, ent["kernel"])
            break
        except Exception:
            pass
    roof["kernel"] = row["name"] + " (" + row["kernel"] + ")"
    roof["kernel_ms"] = row["s"] * 1e3
    roof["share_of_step"] = row["s"] * 1e3 / ms_per_step
    roof["peak_source"] = pk["source"]
    roof["tensor_tflops"], roof["hbm_gbs"] = tf, gb
    roof["timing"] = "isolated launch, CUDA events on the launching stream, L2 flushed, mean of 10"
    if note:
        roof["note"] = note
    return roof


class GpuWorkload:
    """One (workload, events per GPU) configuration on this rank: model, captured step, host + device batches."""

    def __init__(self, args, workload, B, rank, world, dev, n_batches=1):
        from waveformml_b200 import harness, stacks
        from waveformml_b200.synth import make_events
        self.args, self.workload, self.B, self.rank, self.world, self.dev = args, workload, B, rank, world, dev
        full = WORKLOADS[workload][1]
        torch.manual_seed(0)
        self.model = stacks.PSDClassifier().to(dev).train()
        # batch 0 is THE batch of the headline numbers (seed 1234 + rank); the others only feed `rotating`
        self.batches = [make_events(B, n_samples=150, seed=1234 + rank + 1000 * i, full_grid=full) for i in range(n_batches)]
        self.batch = self.batches[0]
        rows_per_event = 154 if full else 10  # multiplicity is clipped to 10 hits/event
        self.capacity = B * rows_per_event
        if args.mode == "graph":
            # two input sets (each with its own captured graph over the same parameters): prefetch() copies the next
            # pinned batch straight into the idle set while the current step runs
            self.step = harness.GraphTrainStep(self.model, "psd", B, self.capacity, 300, n_buffers=2)
        else:
            self.step = harness.TrainStep(self.model, "psd")
        self.host = [tuple(torch.from_numpy(b[k]).pin_memory() for k in ("coords", "wave", "labels")) for b in self.batches]
        self.devb = [tuple(t.to(dev) for t in h) for h in self.host]
        self.rows = int(self.batch["coords"].shape[0])

    def capture(self):
        from waveformml_b200 import harness
        if self.args.mode != "graph":
            return
        self.step.load(*self.devb[0])
        try:
            self.step.capture()
        except Exception as exc:  # e.g. a collective that cannot be captured: keep the update outside the graph
            if self.world == 1:
                raise
            sys.stderr.write("full-step capture failed (%s); capturing forward+backward only\n" % exc)
            self.step = harness.GraphTrainStep(self.model, "psd", self.B, self.capacity, 300, capture_update=False,
                                               n_buffers=2)
            self.step.load(*self.devb[0])
            self.step.capture()

    def preload(self, i=0):
        """Puts batch i into the step's input buffers (graph mode): the inputs are then resident in HBM."""
        if self.args.mode == "graph":
            self.step.load(*self.devb[i % len(self.devb)], buf=0)

    # one step with the batch already resident in HBM.  i = None: the batch preload() put into the step's input
    # buffers (the timed region is the step itself); i = k: batch k is first moved there from another device buffer
    # (one more launch inside the timed region -- the `rotating` figure)
    def step_resident(self, i=None):
        from waveformml_b200 import batcher
        if self.args.mode == "graph":
            if i is not None:
                self.step.load(*self.devb[i % len(self.devb)], buf=0)
            return self.step.run()
        c, w, y = self.devb[(i or 0) % len(self.devb)]
        idx, feats = batcher.pack_batch(c, w)
        return self.step.step(idx, feats, y, self.B)

    def h2d_bytes(self):
        c, w, y = self.host[0]
        return int(c.numel() * c.element_size() + w.numel() * w.element_size() + y.numel() * y.element_size())


def exchange_kind(wl, world):
    if world == 1:
        return "none (one GPU)"
    opt = getattr(wl.step, "opt", None)
    if getattr(opt, "p2p", None) is not None:
        return ("wfsp_sgd_step_p2p -- reduce-scatter in rank order + SGD on the shard + all-gather of the parameters over "
                "NVLink peer memory, one launch")
    return "NCCL all-reduce of gradient buckets issued during backward (head first) + flat SGD"


def timed_steps(fn, steps, flush, sync_all):
    """CUDA events on the launching stream around every step, L2 flushed before each; returns total ms."""
    evs = []
    sync_all()
    for i in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        evs.append((a, b))
    sync_all()
    return sum(a.elapsed_time(b) for a, b in evs)


def time_e2e(wl, steps, flush, sync_all):
    """End to end through the public API: pinned host buffers -> device -> step -> loss on the host.  Graph mode
    uses GraphTrainStep.prefetch (two input sets, each with its own captured graph): the timed region of step i holds
    the H2D copy of batch i+1 straight into the idle set (copy stream, overlapped with the compute of step i), the
    replay and the D2H read of the loss -- one H2D and one D2H per step, K of each over K steps."""
    from waveformml_b200 import batcher
    step, dev = wl.step, wl.dev
    nb = len(wl.host)

    if wl.args.mode == "graph":
        step.prefetch(*wl.host[0])

        def one(i):
            step.run()                             # consumes the pending prefetch
            step.prefetch(*wl.host[(i + 1) % nb])  # next batch: copy stream, overlaps this step
            return step.loss_value()               # D2H of the loss is a node of the step; one stream sync here
    else:
        def one(i):
            c, w, y = (t.to(dev, non_blocking=True) for t in wl.host[i % nb])
            idx, feats = batcher.pack_batch(c, w)
            return float(step.step(idx, feats, y, wl.B).item())

    for i in range(2):
        one(i)
    sync_all()
    total = 0.0
    for i in range(steps):
        flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one(i)
        total += time.perf_counter() - t0
    sync_all()
    if wl.args.mode == "graph" and step._pending is not None:
        step._consume_prefetch()  # leave no batch pending
        torch.cuda.synchronize()
    return total * 1e3


def run_ours(args):
    import torch.distributed as dist
    from waveformml_b200 import _lib, batcher, spconv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    for kv in os.environ.get("WFSP_OPTIONS", "").split(","):  # tuning knobs, e.g. WFSP_OPTIONS=apply_k_split=0
        if kv:
            k, v = kv.split("=")
            _lib.check(lib.wfsp_set_option(k.encode(), int(v)))
    spconv.set_math_mode(args.math)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.zero_()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    B = args.batch
    wl = GpuWorkload(args, args.workload, B, rank, world, dev, n_batches=args.rotate if args.mode == "graph" else 1)
    wl.capture()
    warm = max(args.warmup, 3)
    wl.preload(0)
    for _ in range(warm):
        wl.step_resident()
    sync_all()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.wfsp_kernel_launches()
    dev_ms = timed_steps(lambda i: wl.step_resident(), args.steps, flush, sync_all)
    launches = lib.wfsp_kernel_launches() - launches0
    if args.mode == "graph":  # kernels replayed by the graph
        launches = wl.step.launches_per_replay * args.steps
    e2e_ms = time_e2e(wl, args.steps, flush, sync_all)
    wl.preload(0)

    # sustained: many more steps than the driver's K (the K-step region is a few ms and noise-prone); same
    # per-step event timing with the L2 flush, plus the back-to-back rate without flushes
    extra = {}
    if args.sustained > 0:
        sus_ms = timed_steps(lambda i: wl.step_resident(), args.sustained, flush, sync_all)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.sustained):
            wl.step_resident()
        b.record()
        sync_all()
        bb_ms = a.elapsed_time(b)
        sus_ms, bb_ms = reduce_max([sus_ms, bb_ms])
        extra["sustained"] = {"replays": args.sustained, "ms_per_step": sus_ms / args.sustained,
                              "value": world * B / (sus_ms / args.sustained * 1e-3),
                              "ms_per_step_back_to_back_no_flush": bb_ms / args.sustained,
                              "value_back_to_back_no_flush": world * B / (bb_ms / args.sustained * 1e-3)}
    # rotating: >= 8 DISTINCT batches (different row counts) through the ONE captured graph whose launch shapes were
    # recorded from batch 0 -- what a real epoch pays when the hints do not match the batch
    if args.mode == "graph" and args.rotate > 1:
        rot_steps = max(args.steps, 4 * args.rotate)
        for i in range(args.rotate):
            wl.step_resident(i)
        rot_ms = timed_steps(lambda i: wl.step_resident(i), rot_steps, flush, sync_all)
        (rot_ms,) = reduce_max([rot_ms])
        rws = [int(b["coords"].shape[0]) for b in wl.batches]
        extra["rotating"] = {"batches": args.rotate, "steps": rot_steps, "rows_min": min(rws), "rows_max": max(rws),
                             "hint_rows": rws[0], "ms_per_step": rot_ms / rot_steps,
                             "value": world * B / (rot_ms / rot_steps * 1e-3)}
    # BASELINE.json configs[2] (C3): the z-position regression model (SingleEndedZCNN.json: SparseConv2d(300,150,3,1,1) .
    # BN . ReLU . SparseConv2d(150,1,1) . ReLU . ToDense, masked-L1 segment loss) at 1024 events -- same captured-step
    # machinery, same timing rules
    if args.mode == "graph" and world == 1 and args.c3_batch > 0:
        from waveformml_b200 import harness as _h, stacks as _st
        from waveformml_b200.synth import make_events as _mk
        cb = args.c3_batch
        torch.manual_seed(0)
        zmodel = _st.ZRegressor().to(dev).train()
        zev = _mk(cb, n_samples=150, seed=4321)
        zc, zw, zz = (torch.from_numpy(zev[k]).to(dev) for k in ("coords", "wave", "z"))
        zstep = _h.GraphTrainStep(zmodel, "z", cb, cb * 10, 300)
        zstep.load(zc, zw, zz)
        zstep.capture()
        for _ in range(3):
            zstep.run()
        z_steps = max(10, min(args.steps, 30))
        z_ms = timed_steps(lambda i: zstep.run(), z_steps, flush, sync_all)
        extra["c3"] = {"workload": "C3: z-position regression model (SingleEndedZCNN.json stack), masked-L1 segment loss, "
                                   "synthetic clustered events", "events_per_gpu": cb, "rows_per_gpu": int(zc.shape[0]),
                       "steps": z_steps, "ms_per_step": z_ms / z_steps, "value": cb / (z_ms / z_steps * 1e-3),
                       "unit": "events/s", "gpu_launches_per_step": zstep.launches_per_replay}
        del zstep, zmodel
    # the tight-tolerance mode on the same tensor-core kernels (WFSP_MATH_BF16X3: hi/lo bf16 operand split, fp32-grade
    # results) and, for reference, the CUDA-core fp32 mode it replaces -- same workload, same timing rules
    if args.mode == "graph" and args.math == "bf16" and world == 1 and not args.no_math_modes:
        modes = {}
        for m in ("bf16x3", "fp32"):
            spconv.set_math_mode(m)
            try:
                mw = GpuWorkload(args, args.workload, B, rank, world, dev)
                mw.capture()
                mw.preload(0)
                for _ in range(3):
                    mw.step_resident()
                m_steps = max(10, min(args.steps, 30))
                m_ms = timed_steps(lambda i: mw.step_resident(), m_steps, flush, sync_all)
                (m_ms,) = reduce_max([m_ms])
                modes[m] = {"ms_per_step": m_ms / m_steps, "value": world * B / (m_ms / m_steps * 1e-3)}
                if hasattr(mw.step, "finish"):
                    mw.step.finish()
                del mw
            finally:
                spconv.set_math_mode(args.math)
        modes["bf16x3"]["what"] = "every GEMM operand split into hi + lo bf16, three tcgen05 products per pair, fp32 accumulate"
        modes["fp32"]["what"] = "fp32 FMA on CUDA cores (exact fp32 products)"
        extra["math_modes"] = modes
    clocks = sampler.stop()

    dev_ms, e2e_ms = reduce_max([dev_ms, e2e_ms])
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.math == "bf16" else "f32", "data": "synthetic",
        "config": config_of(args, world, wl.rows),
        "implementation": {
            "parallelism": "dp%d (events sharded by rank, rank-local rulebooks and BatchNorm; gradient exchange inside the "
                           "captured step: %s)" % (world, exchange_kind(wl, world)),
            "math": {"bf16": "bf16 operands / fp32 accumulate (tcgen05)", "fp32": "fp32 CUDA cores",
                     "bf16x3": "fp32-grade on tcgen05: hi/lo bf16 operand split, three products per pair, fp32 "
                               "accumulate"}[args.math],
            "execution": ("whole step replayed from one CUDA graph, row counts on the device, no host readback"
                          if args.mode == "graph" else "eager, exact shapes, one readback per rulebook"),
            "e2e": "two input sets, each with its own captured graph: GraphTrainStep.prefetch copies the next pinned host "
                   "batch straight into the idle set on a copy stream while the current step runs; loss read back "
                   "every step"},
        "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extra)

    pk = peaks()
    if rank == 0 and not args.no_breakdown:
        c, w, _ = wl.devb[0]
        idx, feats = batcher.pack_batch(c, w)
        rows, _ = conv_breakdown(wl.model, idx, feats, B, flush)
        tot = sum(r["s"] for r in rows)
        top = max(rows, key=lambda r: r["s"])
        note = None
        if args.workload == "C2" and B <= 256:
            note = ("%d rows: every launch of this configuration is latency-bound (microseconds of ideal work); the "
                    "throughput-bound configuration is reported under `large`" % wl.rows)
        line["roofline"] = roofline_of(top, pk, ms_per_step, "%s_%d" % (args.workload, B), note)
        line["conv_kernels"] = [{"name": r["name"], "ms": r["s"] * 1e3, "tflops": r["flops"] / r["s"] / 1e12,
                                 "gbs": r["bytes"] / r["s"] / 1e9} for r in rows]
        line["conv_kernels_share_of_step"] = tot * 1e3 / ms_per_step

    # ---- the throughput-bound configuration (C5: full-grid events, 1024 per GPU): the only one of BASELINE.json's
    # configs whose kernels can approach a roofline; same step, same timing rules, reported inside the same line
    if args.large_batch > 0 and not (args.workload == "C5" and B == args.large_batch):
        lw = GpuWorkload(args, "C5", args.large_batch, rank, world, dev)
        lw.capture()
        lw.preload(0)
        for _ in range(3):
            lw.step_resident()
        l_steps = max(10, min(args.steps, 30))
        l_ms = timed_steps(lambda i: lw.step_resident(), l_steps, flush, sync_all)
        l_e2e = time_e2e(lw, l_steps, flush, sync_all)
        l_ms, l_e2e = reduce_max([l_ms, l_e2e])
        lb = args.large_batch
        large = {"workload": WORKLOADS["C5"][0], "events_per_gpu": lb, "global_events_per_step": world * lb,
                 "rows_per_gpu": lw.rows, "steps": l_steps, "ms_per_step": l_ms / l_steps,
                 "value": world * lb / (l_ms / l_steps * 1e-3), "unit": "events/s",
                 "e2e": {"value": world * lb / (l_e2e / l_steps * 1e-3), "unit": "events/s",
                         "h2d_bytes_per_step": lw.h2d_bytes(), "d2h_bytes_per_step": 4},
                 "l2": "inputs and activations exceed L2 (>= 96 MB of waveforms per step); flushed as well"}
        if rank == 0 and not args.no_breakdown:
            c, w, _ = lw.devb[0]
            idx, feats = batcher.pack_batch(c, w)
            rows, _ = conv_breakdown(lw.model, idx, feats, lb, flush)
            top = max(rows, key=lambda r: r["s"])
            large["roofline"] = roofline_of(top, pk, l_ms / l_steps, "C5_%d" % lb)
            large["conv_kernels"] = [
                {"name": r["name"], "ms": r["s"] * 1e3, "tflops": r["flops"] / r["s"] / 1e12, "gbs": r["bytes"] / r["s"] / 1e9,
                 "frac_burst": r["flops"] / r["s"] / 1e12 / pk["tflops_burst"], "frac_hbm": r["bytes"] / r["s"] / 1e9 / pk["hbm_gbs"]}
                for r in rows]
            large["conv_kernels_share_of_step"] = sum(r["s"] for r in rows) * 1e3 / (l_ms / l_steps)
            # ideal step time of this configuration: max of tensor and HBM time
            fl = sum(r["flops"] for r in rows)
            by = sum(r["bytes"] for r in rows)
            ideal_ms = max(fl / (pk["tflops_burst"] * 1e12), by / (pk["hbm_gbs"] * 1e9)) * 1e3
            large["step_vs_ideal"] = {"conv_flops": fl, "conv_bytes": by, "ideal_ms": ideal_ms,
                                      "frac": ideal_ms / (l_ms / l_steps)}
            # the HBM-bound passes between the convolutions, at the sizes of this workload's first two blocks
            n0 = lw.rows  # full-grid events: 154 rows in, 108 rows out of the first 3x3 layer per event
            large["hbm_kernels"] = bn_breakdown([("L0", n0, 252), ("L1", n0 * 108 // 154, 158)], flush, pk)
        line["large"] = large
    wl_cpu_model, wl_cpu_batch = wl.model, wl.batch

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cstep = cpu_step_fn(cpu_reference_model(wl_cpu_model), wl_cpu_batch, wl_cpu_model.n_linear)
        pick_cpu_threads(cstep)
        times = time_cpu(cstep, args.cpu_seconds)
        cms = float(np.mean(times))
        line["cpu_baseline"] = {"value": B / cms, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d full training steps of the same %d-event batch (restated spconv-1.2.1 "
                                          "CPU algorithm, oracle/)" % (len(times), B)}
    if rank == 0:
        emit(line)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # The captured graph holds NCCL kernels; tearing the communicator down under it hung the 2-GPU run
        # at exit (the JSON line was already out).  Drain the device, meet the other ranks once, and leave
        # without NCCL teardown.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)
    return 0


def _quiet_stdout():
    """Libraries print to fd 1 behind our back (NCCL's version banner, cuBLAS notices).  The contract is ONE
    JSON line on stdout: send fd 1 to stderr for the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    a = parse()
    _RESULT = _quiet_stdout()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
