#!/bin/bash
# usage: gpu_scale.sh TAG N [extra bench args]: one torchrun bench at N GPUs, JSON to gpurun_out/TAG_Ngpu.json
tag=$1; n=$2; shift 2
port=$((29600 + n))
if [ "$n" = "1" ]; then
  timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --large-batch 0 --no-breakdown --no-cpu-baseline --no-math-modes --c3-batch 0 --rotate 1 --sustained 0 "$@" > gpurun_out/${tag}_${n}gpu.json 2> gpurun_out/${tag}_${n}gpu.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 50 --warmup 5 --large-batch 0 --no-breakdown --no-cpu-baseline --no-math-modes --c3-batch 0 --rotate 1 --sustained 0 "$@" > gpurun_out/${tag}_${n}gpu.json 2> gpurun_out/${tag}_${n}gpu.err
fi
python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_${n}gpu.json'))
print('${tag}', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'])" || tail -5 gpurun_out/${tag}_${n}gpu.err
