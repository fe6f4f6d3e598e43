#!/bin/bash
set -u
mkdir -p gpurun_out
for ph in 0 1; do
  echo "== phased=$ph pytest"; BENCHPUSH_BD_PHASED=$ph timeout 900 python -m pytest tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py -x -q 2>&1 | tail -3
  echo "== phased=$ph area-clearing 4096"; BENCHPUSH_BD_PHASED=$ph timeout 300 python tools/areaclr_time.py 4096 4 2 2>&1 | tail -2
  echo "== phased=$ph area-clearing 16384"; BENCHPUSH_BD_PHASED=$ph timeout 300 python tools/areaclr_time.py 16384 3 2 2>&1 | tail -2
  echo "== phased=$ph box-delivery 4096"; BENCHPUSH_BD_PHASED=$ph timeout 300 python tools/boxdel_time.py 4096 6 2 2>&1 | tail -3 | head -2
done 2>&1 | tee gpurun_out/variants_phased.log
