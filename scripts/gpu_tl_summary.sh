#!/bin/bash
# usage: gpu_tl_summary.sh TAG "OPTIONS" [WORKLOAD BATCH]: timeline with options, print span + conv kernel durations
tag=$1; opts=$2; wl=${3:-C2}; b=${4:-64}
WFSP_OPTIONS="$opts" timeout 200 python scripts/gpu_timeline.py $wl $b > gpurun_out/tl_${tag}.txt 2>&1
echo "== $tag ($opts): $(grep 'step span' gpurun_out/tl_${tag}.txt)"
grep -E "conv_apply|conv_wgrad" gpurun_out/tl_${tag}.txt | awk '{printf "%s/%s ", $1, $2}'; echo
