#!/bin/bash
# One full GPU-box visit: parity tests, smoke, bench lines for the three workloads (with CPU baseline), the
# reference arm, the kernel timeline, and (unless "noncu") the ncu launch lists + full captures.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh tests/test_gpu_graph.py 2>&1 | grep -E "^==|passed|failed|rror"
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
for cfg in "C2 64" "C2 1024" "C5 1024"; do
  set -- $cfg
  timeout 900 python bench.py --workload $1 --batch $2 --steps 30 --warmup 5 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "== bench $1 batch $2 rc=$?"; cut -c1-330 gpurun_out/bench_$1_$2.json; tail -n 3 gpurun_out/bench_$1_$2.err
done
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_C2_64.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref_C2_64.json
timeout 300 python scripts/gpu_timeline.py C2 64 > gpurun_out/timeline_C2_64.txt 2>&1
timeout 300 python scripts/gpu_timeline.py C5 1024 > gpurun_out/timeline_C5_1024.txt 2>&1
head -4 gpurun_out/timeline_C2_64.txt | tail -1; head -4 gpurun_out/timeline_C5_1024.txt | tail -1
if [ "$1" != "noncu" ] && [ "$NONCU" != "1" ]; then
  timeout 500 bash scripts/gpu_profile.sh final_C2_64 C2 64
  timeout 500 bash scripts/gpu_profile.sh final_C5_1024 C5 1024
fi
