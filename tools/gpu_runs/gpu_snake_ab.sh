#!/bin/bash
# A/B of the snake-order dealing in the balanced step schedule (BP_NO_SNAKE=1: plain round robin); state hashes must agree.
mkdir -p gpurun_out; : > gpurun_out/r02h_snake_ab.log
timeout 900 python -m pytest tests/test_gpu_shipice.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/r02h_snake_ab.log
for E in 8192 16384 4096 32768; do
  for ns in 1 0; do
    echo "E=$E BP_NO_SNAKE=$ns" >> gpurun_out/r02h_snake_ab.log
    BP_NO_SNAKE=$ns timeout 900 python tools/bench_variant.py $E 0.5 3 6 34 >> gpurun_out/r02h_snake_ab.log 2>&1
  done
done
cat gpurun_out/r02h_snake_ab.log
