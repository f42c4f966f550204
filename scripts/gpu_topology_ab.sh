#!/bin/bash
# A/B of graph-topology choices on the 64-event step: tests, then a timeline and a bench line per variant.
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python scripts/gpu_timeline.py C2 64 > gpurun_out/topo_${name}_timeline.txt 2>&1
  env "$@" timeout 300 python bench.py --large-batch 0 --c3-batch 0 --no-cpu-baseline --no-math-modes --no-breakdown --steps 50 --warmup 5 > gpurun_out/topo_${name}_bench.json 2> gpurun_out/topo_${name}_bench.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open("gpurun_out/topo_%s_bench.json" % n).read().strip().splitlines()[-1])
tl = open("gpurun_out/topo_%s_timeline.txt" % n).read().splitlines()
conv1 = [l for l in tl if "conv_apply_umma_kernel" in l][:1]
span = [l for l in tl if l.startswith("step span")]
print(n, "%.4f ms" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "|", span[0] if span else "", "| conv1 at", conv1[0].split()[0] if conv1 else "?")
PY
}
run A WFSP_X=0
run B WFSP_OPTIONS=prep_ctas=96
run C WFSP_OPTIONS=prep_ctas=64
run B2 WFSP_OPTIONS=prep_ctas=96
