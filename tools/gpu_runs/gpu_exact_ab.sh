#!/bin/bash
# Room envs + maze on one B200: GPU tests, then bench lines with the default reference-order spfa parents and with --fast-parents.
mkdir -p gpurun_out; : > gpurun_out/r02_exact_ab.log
timeout 900 python -m pytest tests/test_gpu_spfa.py tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py tests/test_gpu_maze.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_exact_ab.log
timeout 900 python bench.py --env maze --steps 10 --warmup 3 > gpurun_out/r02d_bench_maze.json 2> gpurun_out/r02d_bench_maze.err; tail -2 gpurun_out/r02d_bench_maze.err >> gpurun_out/r02_exact_ab.log
cut -c1-300 gpurun_out/r02d_bench_maze.json >> gpurun_out/r02_exact_ab.log
for args in "--env box-delivery" "--env box-delivery --fast-parents" "--env area-clearing" "--env area-clearing --fast-parents" \
            "--env box-delivery --obstacle-config small_columns" "--env box-delivery --obstacle-config small_columns --fast-parents" \
            "--env box-delivery --obstacle-config large_columns" "--env box-delivery --obstacle-config large_columns --fast-parents" \
            "--env box-delivery --obstacle-config large_divider" "--env box-delivery --obstacle-config large_divider --fast-parents"; do
  echo "bench.py $args" >> gpurun_out/r02_exact_ab.log
  tag=$(echo "$args" | tr -d ' -' )
  timeout 900 python bench.py $args --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/r02d_bench_$tag.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02d_bench_$tag.json').read())
print(round(d['value']), 'env-steps/s  ms/step', round(d['ms_per_step'],1), 'move', round(d['roofline']['kernel_ms'],1), 'obs', round(d['roofline']['obs_kernel_ms'],1), 'e2e', round(d['e2e']['value']), 'status', d['config']['env_status_bits'])" >> gpurun_out/r02_exact_ab.log 2>&1
done
cat gpurun_out/r02_exact_ab.log
