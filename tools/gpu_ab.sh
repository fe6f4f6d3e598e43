#!/bin/bash
# A/B: batched step kernel vs one-env-per-warp (development aid)
mkdir -p gpurun_out; : > gpurun_out/ab.log
E=${1:-16384}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for impl in batch warp; do echo "impl=$impl" >> gpurun_out/ab.log; BP_STEP_IMPL=$impl timeout 600 python tools/bench_variant.py $E 0.5 6 4 >> gpurun_out/ab.log 2>&1; done
cat gpurun_out/ab.log
