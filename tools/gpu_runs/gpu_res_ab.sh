#!/bin/bash
# A/B on the GPU box: resident step (sim_res.cuh) vs batched step (sim_batch.cuh); same state checksum expected.
mkdir -p gpurun_out; : > gpurun_out/res_ab.log
E=${1:-16384}; W=${2:-3}; K=${3:-4}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> gpurun_out/res_ab.log
timeout 900 python -m pytest tests/test_gpu_shipice.py -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/res_ab.log
for impl in res batch; do
  echo "impl=$impl" >> gpurun_out/res_ab.log
  BP_STEP_IMPL=$impl timeout 600 python tools/bench_variant.py $E 0.5 $W $K 34 >> gpurun_out/res_ab.log 2>&1
done
for b in 1 2; do
  echo "impl=res blocks_per_sm=$b" >> gpurun_out/res_ab.log
  BP_RES_BLOCKS_PER_SM=$b BP_STEP_IMPL=res timeout 600 python tools/bench_variant.py $E 0.5 $W $K 34 >> gpurun_out/res_ab.log 2>&1
done
cat gpurun_out/res_ab.log
