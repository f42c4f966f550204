#!/bin/bash
# ncu launch list (per-launch durations) of one short bench run.  Usage: gpu_launchlist.sh TAG WORKLOAD BATCH
TAG=$1; WL=$2; B=$3
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --batch $B --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
