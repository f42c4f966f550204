#!/bin/bash
# Round-2 evidence in one GPU-box visit: smoke, both bench arms, timelines, ncu launch lists and full captures
# (FAST=1: without the --set full captures).
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "== bench default rc=$?"; cut -c1-250 gpurun_out/r2_final_bench.json
timeout 400 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
echo "== reference arm rc=$?"; cut -c1-200 gpurun_out/r2_final_bench_reference.json
timeout 400 python bench.py --workload C2 --batch 1024 --large-batch 0 --no-cpu-baseline --no-math-modes --steps 30 --warmup 5 > gpurun_out/r2_final_bench_C2_1024.json 2>/dev/null
echo "== C2@1024 rc=$?"; cut -c1-200 gpurun_out/r2_final_bench_C2_1024.json
timeout 200 python scripts/gpu_timeline.py C2 64 > gpurun_out/r2_final_timeline_C2_b64.txt 2>&1
timeout 200 python scripts/gpu_timeline.py C5 1024 > gpurun_out/r2_final_timeline_C5_b1024.txt 2>&1
grep "step span" gpurun_out/r2_final_timeline_C2_b64.txt gpurun_out/r2_final_timeline_C5_b1024.txt
for cfg in "C2 64" "C5 1024"; do
  set -- $cfg
  TAG=r2_final_$1_b$2
  CMD="python bench.py --workload $1 --batch $2 --large-batch 0 --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --no-math-modes --c3-batch 0 --rotate 1 --sustained 0"
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/${TAG}_plain.log; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
  echo "launch list $TAG rc=$?"
  if [ "$FAST" != "1" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_(apply|wgrad)" -c 8 -o gpurun_out/${TAG}_conv_full -f $CMD > gpurun_out/${TAG}_ncu_conv.log 2>&1
    echo "full conv $TAG rc=$?"
  fi
done
[ "$FAST" = "1" ] && exit 0  # (the --set full captures are only repeated when a conv / BatchNorm kernel changed)
CMD="python bench.py --workload C5 --batch 1024 --large-batch 0 --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --no-math-modes --c3-batch 0 --rotate 1 --sustained 0"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:bn_stream" -c 9 -o gpurun_out/r2_final_C5_b1024_bn_full -f $CMD > gpurun_out/r2_final_C5_b1024_ncu_bn.log 2>&1
echo "full bn rc=$?"
ls -la gpurun_out/r2_final_* | awk '{print $5, $9}'
