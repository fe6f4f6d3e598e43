#!/bin/bash
# Variant sweep on the GPU box: bash tools/gpu_sweep.sh ENVS "variant files..."
set -u
mkdir -p gpurun_out
E=${1:-16384}; shift
: > gpurun_out/sweep.log
for so in "$@"; do
  BENCHPUSH_B200_LIB=$PWD/$so timeout 600 python tools/bench_variant.py $E 0.5 6 4 >> gpurun_out/sweep.log 2>&1
done
cat gpurun_out/sweep.log
