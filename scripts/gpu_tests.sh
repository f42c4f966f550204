#!/bin/bash
# Runs the GPU parity tests file by file (each under its own timeout so a hung kernel cannot eat the
# whole gpurun slot) and keeps the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_rulebook.py tests/test_gpu_dense_pack.py tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_tma.py tests/test_gpu_scn.py tests/test_window_edges.py "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -x --timeout 300 > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f rc=$r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
