#!/bin/bash
# A/B of the warp-per-env kernels' occupancy (BD_MOVE_MINB = resident 128-thread blocks per SM): env-steps/s per variant
set -u
mkdir -p gpurun_out
for m in 4 5 6 7 8 10; do
  lib=benchpush_b200/csrc/variants/lib_minb$m.so
  [ $m = 4 ] && lib=benchpush_b200/csrc/libbenchpush_b200.so
  echo "== MINB=$m area-clearing 4096"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/areaclr_time.py 4096 4 2 2>&1 | tail -2
  echo "== MINB=$m area-clearing 16384"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/areaclr_time.py 16384 3 2 2>&1 | tail -2
  echo "== MINB=$m box-delivery 4096"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/boxdel_time.py 4096 6 2 2>&1 | tail -3 | head -2
done 2>&1 | tee gpurun_out/variants_minb.log
