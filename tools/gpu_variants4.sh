#!/bin/bash
# A/B of launch-shape variants of the ship-ice step kernel at steady state (development aid): blocks per SM 5 (96 registers),
# 6 and 12 envs per warp, against the default (8 envs per warp, 4 blocks per SM).
set -u
mkdir -p gpurun_out
for v in main k8_b5 k6_b4 k12_b4; do
  lib=benchpush_b200/csrc/variants/lib_$v.so
  [ $v = main ] && lib=benchpush_b200/csrc/libbenchpush_b200.so
  for n in 18944 65536; do
    BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/bench_variant.py $n 0.5 2 3 34 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/variants_r01g.log
