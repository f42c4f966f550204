#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share.  Usage: launch_summary.py launches.csv [last_n_launches | until=SUBSTR]
(until=SUBSTR keeps the launches before the first kernel whose name contains SUBSTR, e.g. the bench's C3 block)."""
import csv, sys, collections, io
path = sys.argv[1]
lines = [l for l in open(path, errors="replace") if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
if len(sys.argv) > 2 and sys.argv[2].startswith("until="):
    cut = next((i for i, r in enumerate(rows) if sys.argv[2][6:] in r["Kernel Name"]), len(rows))
    rows = rows[:cut]
elif len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
def dur_us(r):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        return 0.0
    u = r["Metric Unit"]
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += dur_us(r)
tot = sum(a[1] for a in agg.values())
print("launches %d  total %.1f us" % (len(rows), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6d %10.1f us %5.1f%%  %s" % (a[0], a[1], 100 * a[1] / tot, k))
