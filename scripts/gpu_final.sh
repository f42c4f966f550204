#!/bin/bash
# Round-end evidence in one GPU-box visit: smoke, the headline bench line (with the CPU baseline), the large
# workloads, the reference arm, the kernel timeline, then the ncu launch lists + full captures.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_C2_64.json 2> gpurun_out/bench_C2_64.err
echo "== bench default rc=$?"; cut -c1-300 gpurun_out/bench_C2_64.json; tail -n 2 gpurun_out/bench_C2_64.err
for cfg in "C2 1024" "C5 1024"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "== bench $1 batch $2 rc=$?"; cut -c1-200 gpurun_out/bench_$1_$2.json
done
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_C2_64.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref_C2_64.json
timeout 200 python scripts/gpu_timeline.py C2 64 > gpurun_out/timeline_C2_64.txt 2>&1
timeout 200 python scripts/gpu_timeline.py C5 1024 > gpurun_out/timeline_C5_1024.txt 2>&1
head -4 gpurun_out/timeline_C2_64.txt | tail -1; head -4 gpurun_out/timeline_C5_1024.txt | tail -1
timeout 300 bash scripts/gpu_profile.sh final_C2_64 C2 64
timeout 400 bash scripts/gpu_profile.sh final_C5_1024 C5 1024
