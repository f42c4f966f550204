#!/usr/bin/env python
"""Extracts per-launch DRAM traffic and duration of the conv kernels from an `ncu --set full` report of one
step (captured with -k regex:conv_ -c 8: the first eager pass launches fwd L0, L1, L2, then wgrad L2,
dgrad L2, wgrad L1, dgrad L1, wgrad L0) and prints / merges them into profiles/r1_traffic.json.
Usage: ncu_traffic.py report.ncu-rep KEY [OUT.json]   (KEY e.g. C2_64; OUT defaults to profiles/r2_traffic.json).
With KEY ending in "_bn" the rows are labelled by kernel name + index instead of the conv launch order."""
import csv, io, json, os, subprocess, sys
rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[col[name]].replace(",", "")); u = units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "%": 1, "": 1}.get(u, 1)
    return v * scale
order = ["L0 fwd", "L1 fwd", "L2 fwd", "L2 wgrad", "L2 dgrad", "L1 wgrad", "L1 dgrad", "L0 wgrad"]
res = {}
generic = key.endswith("_bn")
for i, r in enumerate(rows[2:]):
    name = order[i] if (i < len(order) and not generic) else "launch%d" % i
    res[name] = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", ""),
                 "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                 "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
                 "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "grid": r[col["launch__grid_size"]], "regs": r[col["launch__registers_per_thread"]]}
    print("%-9s %-28s %8.1f us  dram %8.2f MB  tensor %5.1f%%  dram %5.1f%%  grid %s" % (
        name, res[name]["kernel"][:28], res[name]["duration_us_under_ncu"], res[name]["dram_bytes"] / 1e6,
        res[name]["tensor_pipe_pct"], res[name]["dram_pct"], res[name]["grid"]))
path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "r2_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = res
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
