#!/bin/bash
# A/B of the k_bd_dist launch shape in the reference-order-parents mode (BP_BD_DIST_WIDE=1: 256-thread blocks as before).
mkdir -p gpurun_out; : > gpurun_out/r02h_dist_ab.log
timeout 600 python -m pytest tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/r02h_dist_ab.log
for args in "--env box-delivery" "--env area-clearing" "--env box-delivery --obstacle-config small_columns" "--env box-delivery --obstacle-config large_divider"; do
  for wide in 0 1; do
    echo "bench.py $args  BP_BD_DIST_WIDE=$wide" >> gpurun_out/r02h_dist_ab.log
    BP_BD_DIST_WIDE=$wide timeout 900 python bench.py $args --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ms=d['ms_per_step']; mv=d['roofline']['kernel_ms']; ob=d['roofline']['obs_kernel_ms']
print(round(d['value']), 'env-steps/s  ms/step', round(ms,1), 'move', round(mv,1), 'obs', round(ob,1), 'rest (begin + dist + finish)', round(ms-mv-ob,1))" >> gpurun_out/r02h_dist_ab.log 2>&1
  done
done
cat gpurun_out/r02h_dist_ab.log
