#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines for the three workloads, ncu launch list.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
for cfg in "C2 64" "C2 1024" "C5 1024"; do
  set -- $cfg
  echo "== bench $1 batch $2"
  timeout 900 python bench.py --workload $1 --batch $2 --steps 20 --warmup 5 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "rc=$?"; tail -c 3000 gpurun_out/bench_$1_$2.json; tail -n 5 gpurun_out/bench_$1_$2.err
done
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_C2_64.json 2>&1; tail -c 1500 gpurun_out/bench_ref_C2_64.json
if [ "$1" != "noncu" ]; then
  echo "== ncu launch list (C2 batch 64)"
  CMD="python bench.py --workload C2 --batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown"
  $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C2_64.csv $CMD > gpurun_out/ncu_run.log 2>&1
  echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_run.log
fi
