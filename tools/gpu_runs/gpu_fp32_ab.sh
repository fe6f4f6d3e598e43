#!/bin/bash
# fp64 vs fp32-state build of the ship-ice step (batched and resident implementations) at the bench's batch sizes.
mkdir -p gpurun_out; : > gpurun_out/fp32_ab.log
E=${1:-65536}
for prec in fp64 fp32; do
  for impl in batch res; do
    echo "precision=$prec impl=$impl E=$E" >> gpurun_out/fp32_ab.log
    BP_PRECISION=$prec BP_STEP_IMPL=$impl timeout 900 python tools/bench_variant.py $E 0.5 3 5 34 >> gpurun_out/fp32_ab.log 2>&1
  done
done
cat gpurun_out/fp32_ab.log
