#!/bin/bash
# Cost of the reference-order spfa parents (exact_spfa_parents) at the bench sizes of the room envs.
mkdir -p gpurun_out; : > gpurun_out/r02_exact_ab.log
timeout 600 python -m pytest tests/test_gpu_spfa.py tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_exact_ab.log
for args in "--env box-delivery" "--env box-delivery --exact-parents" "--env area-clearing" "--env area-clearing --exact-parents" \
            "--env box-delivery --obstacle-config small_columns" "--env box-delivery --obstacle-config small_columns --exact-parents" \
            "--env box-delivery --obstacle-config large_divider" "--env box-delivery --obstacle-config large_divider --exact-parents"; do
  echo "bench.py $args" >> gpurun_out/r02_exact_ab.log
  timeout 900 python bench.py $args --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), 'env-steps/s  ms/step', round(d['ms_per_step'],1), 'move', round(d['roofline']['kernel_ms'],1), 'obs', round(d['roofline']['obs_kernel_ms'],1), 'e2e', round(d['e2e']['value']), 'status', d['config']['env_status_bits'])" >> gpurun_out/r02_exact_ab.log 2>&1
done
cat gpurun_out/r02_exact_ab.log
