#!/bin/bash
# Quick GPU visit while iterating on kernels: parity tests, then short bench lines (no CPU baseline).
# Usage: gpu_quick.sh [notests] -- prints a compact per-kernel summary for each workload.
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  bash scripts/gpu_tests.sh tests/test_gpu_graph.py 2>&1 | grep -E "^==|passed|failed|error|Error|assert" | head -40
fi
for cfg in "C2 64" "C2 1024" "C5 1024"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --batch $2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_$1_$2.json 2> gpurun_out/q_$1_$2.err
  echo "rc=$?"; tail -n 5 gpurun_out/q_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
wl, b = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/q_%s_%s.json" % (wl, b)).read().strip().splitlines()[-1])
    print(wl, b, "%.3f ms" % d["ms_per_step"], "%.0f ev/s" % d["value"], "e2e %.0f" % d["e2e"]["value"],
          "launches/step", d["gpu_launches"] / d["steps"], "roof", d.get("roofline", {}).get("frac"))
    for k in d.get("conv_kernels", []):
        print("    %-28s %7.1f us %6.1f TF %6.0f GB/s" % (k["name"], k["ms"] * 1e3, k["tflops"], k["gbs"]))
except Exception as e:
    print("no line:", e)
PY
done
