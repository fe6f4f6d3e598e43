#!/bin/bash
# Scaling points of the headline bench on one box: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_runs/gpu_scale_r02.sh N [N2 ...]'
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_scale_1gpu.json 2> gpurun_out/r02_scale_1gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err
  fi
  tail -n 1 gpurun_out/r02_scale_${N}gpu.json | cut -c1-400
done
