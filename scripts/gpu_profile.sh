#!/bin/bash
# ncu evidence for one workload: launch list of the whole bench command (per-launch durations) and a
# `--set full` capture of the conv kernels of one step.  Usage: gpu_profile.sh TAG WORKLOAD BATCH [KREGEX] [COUNT]
# Every ncu pass only runs after the same command exited 0 without ncu (B200_PROFILING.md).
TAG=$1; WL=$2; B=$3; KRE=${4:-"regex:conv_(apply|wgrad)"}; CNT=${5:-8}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --batch $B --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KRE" -c $CNT \
  -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_*
