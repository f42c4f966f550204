#!/usr/bin/env python
"""Per-source-line warp-stall samples of each kernel launch in an .ncu-rep captured with
--import-source on (library built with -lineinfo).  Usage: ncu_lines.py report.ncu-rep [launch_index ...]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
sel = [int(x) for x in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
launches, cur, hdr, cur_file = [], None, None, None
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]
        # one section per (launch, source file): a file seen before starts the next launch
        if cur is None or cur_file in cur["files"]:
            cur = {"name": "", "lines": collections.OrderedDict(), "exec": collections.Counter(), "files": set()}
            launches.append(cur)
        cur["files"].add(cur_file); hdr = None; continue
    if len(r) >= 2 and r[0] == "Function Name":
        cur["name"] = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if cur is None or hdr is None or len(r) != len(hdr):
        continue
    if r[0]:  # a source-line row carries the aggregate of the SASS rows below it
        line = (cur_file.split("/")[-1], r[0], r[1])
        cur["lines"][line] = cur["lines"].get(line, 0) + int(r[si] or 0 if r[si] != "-" else 0)
        cur["exec"][line] += int(r[ei] or 0 if r[ei] != "-" else 0)
for i, L in enumerate(launches):
    if sel and i not in sel:
        continue
    tot = sum(L["lines"].values()); te = sum(L["exec"].values())
    print("== launch %d %s  samples %d  warp-instr %d" % (i, L["name"][:70], tot, te))
    for (f, n, src), v in L["lines"].items():
        if tot and v >= 0.012 * tot:
            print("  %5d %5.1f%% exec %4.1f%%  %s:%s  %s" % (v, 100.0 * v / tot, 100.0 * L["exec"][(f, n, src)] / max(te, 1), f, n, src.strip()[:100]))
