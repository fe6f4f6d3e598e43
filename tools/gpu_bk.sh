#!/bin/bash
# envs per warp of the ship-ice step kernel at the per-GPU load of the 8-GPU strong-scaling point (8192 envs) and of 4 GPUs (16384)
set -u
mkdir -p gpurun_out
for v in main bk4 bk2; do
  lib=benchpush_b200/csrc/variants/lib_$v.so
  [ $v = main ] && lib=benchpush_b200/csrc/libbenchpush_b200.so
  for n in 8192 16384; do
    echo "== $v envs=$n"; BENCHPUSH_B200_LIB=$PWD/$lib timeout 600 python bench.py --envs $n --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('value %.0f ms/step %.1f kernel_ms %.1f'%(l['value'],l['ms_per_step'],l['roofline']['kernel_ms']))"
  done
done 2>&1 | tee gpurun_out/variants_bk.log
