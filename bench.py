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
    ap.add_argument("--math", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: sync-free capacity-sized step replayed from a CUDA graph; eager: exact shapes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


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
        "config": {"workload": WORKLOADS[args.workload][0], "events_per_step": args.batch,
                   "rows": int(batch["coords"].shape[0])},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d full training steps of %d events (restated spconv-1.2.1 CPU algorithm: C hash "
                                   "rulebook + per-offset gather/torch.mm/scatter-add, torch-CPU BN/ReLU/Linear/SGD); "
                                   "thread count = fastest of {1,4,16,64,all} on this host (%d cores)"
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


def run_ours(args):
    import torch.distributed as dist
    from waveformml_b200 import _lib, batcher, harness, spconv, stacks

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
    spconv.set_math_mode(args.math)
    torch.manual_seed(0)
    model = stacks.PSDClassifier().to(dev).train()
    batch = make_batch(args, rank)
    B = args.batch
    rows_per_event = 154 if WORKLOADS[args.workload][1] else 10  # multiplicity is clipped to 10 hits/event
    if args.mode == "graph":
        step = harness.GraphTrainStep(model, "psd", B, B * rows_per_event, 300)
    else:
        step = harness.TrainStep(model, "psd")
    h_coords = torch.from_numpy(batch["coords"]).pin_memory()
    h_wave = torch.from_numpy(batch["wave"]).pin_memory()
    h_labels = torch.from_numpy(batch["labels"]).pin_memory()
    d_coords, d_wave, d_labels = h_coords.to(dev), h_wave.to(dev), h_labels.to(dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.zero_()

    if args.mode == "graph":
        def step_resident():
            step.load(d_coords, d_wave, d_labels)  # device-to-device into the graph's static buffers
            return step.run()

        def step_e2e():
            step.load(h_coords, h_wave, h_labels)  # async H2D from pinned memory
            return float(step.run().item())

        step.load(d_coords, d_wave, d_labels)
        try:
            step.capture()
        except Exception as exc:  # e.g. a collective that cannot be captured: keep the update outside the graph
            if world == 1:
                raise
            sys.stderr.write("full-step capture failed (%s); capturing forward+backward only\n" % exc)
            step = harness.GraphTrainStep(model, "psd", B, B * rows_per_event, 300, capture_update=False)
            step.load(d_coords, d_wave, d_labels)
            step.capture()
        n0 = lib.wfsp_kernel_launches()
    else:
        def step_resident():
            idx, feats = batcher.pack_batch(d_coords, d_wave)
            return step.step(idx, feats, d_labels, B)

        def step_e2e():
            c = h_coords.to(dev, non_blocking=True)
            w = h_wave.to(dev, non_blocking=True)
            y = h_labels.to(dev, non_blocking=True)
            idx, feats = batcher.pack_batch(c, w)
            return float(step.step(idx, feats, y, B).item())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.wfsp_kernel_launches()
    evs = []
    sync_all()
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_resident()
        b.record()
        evs.append((a, b))
    sync_all()
    launches = lib.wfsp_kernel_launches() - launches0
    if args.mode == "graph":  # kernels replayed by the graph: count what one captured step enqueued
        launches = step.launches_per_replay * args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # end to end: pinned host buffers -> device -> step -> loss on the host
    for _ in range(2):
        step_e2e()
    sync_all()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
    sync_all()
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.math == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][0], "events_per_gpu": B, "global_events_per_step": world * B,
                   "rows_per_gpu": int(d_coords.shape[0]), "parallelism": "dp%d (events sharded by rank, NCCL "
                   "all-reduce of the flat 4.2 MB gradient)" % world, "l2": "flushed with a 256 MB write before every timed step",
                   "math": "bf16 operands / fp32 accumulate (tcgen05)" if args.math == "bf16" else "fp32 CUDA cores",
                   "execution": ("whole step replayed from one CUDA graph, row counts on the device, no host readback"
                                 if args.mode == "graph" else "eager, exact shapes, one readback per rulebook")},
        "e2e": {"value": e2e_value, "unit": "events/s",
                "h2d_bytes_per_step": int(h_coords.numel() * 4 + h_wave.numel() * 2 + h_labels.numel() * 8),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks,
    }

    if rank == 0 and not args.no_breakdown:
        pk = peaks()
        idx, feats = batcher.pack_batch(d_coords, d_wave)
        rows, _ = conv_breakdown(model, idx, feats, B, flush)
        tot = sum(r["s"] for r in rows)
        top = max(rows, key=lambda r: r["s"])
        ai = top["flops"] / top["bytes"]
        ridge = pk["tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        if ai >= ridge:
            roof = {"bound": "tensor", "achieved": top["flops"] / top["s"] / 1e12, "peak": pk["tflops_sustained"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": top["bytes"] / top["s"] / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        # DRAM bytes of the same launch from the committed `ncu --set full` capture of this command line
        # (profiles/r1_traffic.json, made by scripts/ncu_traffic.py); null if that capture does not exist
        roof["traffic"] = None
        roof["algorithmic_bytes"] = top["bytes"]
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            parts = top["name"].split()
            ent = tr["%s_%d" % (args.workload, B)][parts[0] + " " + parts[-1]]
            roof["traffic"] = ent["dram_bytes"]
            roof["traffic_source"] = "profiles/r1_traffic.json (ncu --set full, %s)" % ent["kernel"]
        except Exception:
            pass
        roof["kernel"] = top["name"] + " (" + top["kernel"] + ")"
        roof["kernel_ms"] = top["s"] * 1e3
        roof["share_of_step"] = top["s"] * 1e3 / ms_per_step
        roof["peak_source"] = pk["source"]
        roof["tensor_tflops"] = top["flops"] / top["s"] / 1e12
        roof["hbm_gbs"] = top["bytes"] / top["s"] / 1e9
        line["roofline"] = roof
        line["conv_kernels"] = [{"name": r["name"], "ms": r["s"] * 1e3, "tflops": r["flops"] / r["s"] / 1e12,
                                 "gbs": r["bytes"] / r["s"] / 1e9} for r in rows]
        line["conv_kernels_share_of_step"] = tot * 1e3 / ms_per_step

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cstep = cpu_step_fn(cpu_reference_model(model), batch, model.n_linear)
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
