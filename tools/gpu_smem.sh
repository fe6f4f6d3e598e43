#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest room envs"; timeout 900 python -m pytest tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py -x -q 2>&1 | tail -3
for v in main smem_minb6 smem_minb7; do
  lib=benchpush_b200/csrc/variants/lib_$v.so
  [ $v = main ] && lib=benchpush_b200/csrc/libbenchpush_b200.so
  echo "== $v area-clearing 4096"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/areaclr_time.py 4096 4 2 2>&1 | tail -2
  echo "== $v area-clearing 16384"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/areaclr_time.py 16384 3 2 2>&1 | tail -2
  echo "== $v box-delivery 4096"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 300 python tools/boxdel_time.py 4096 6 2 2>&1 | tail -3 | head -2
done 2>&1 | tee gpurun_out/variants_smem.log
