#!/bin/bash
# Round-2 final measurement set on one B200: GPU test suite, smoke, headline bench + reference arm, secondary env benches,
# ncu launch list of the headline command.   gpurun --timeout 3000 -- 'bash tools/gpu_runs/gpu_final_r02.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02f_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02f_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02f_gputest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -1 gpurun_out/r02f_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02f_bench_ref.json 2>> gpurun_out/r02f_bench.err
timeout 600 python bench.py --env box-delivery --steps 6 --warmup 3 > gpurun_out/r02f_bench_boxdel.json 2>> gpurun_out/r02f_bench.err
timeout 600 python bench.py --env area-clearing --steps 6 --warmup 3 > gpurun_out/r02f_bench_areaclr.json 2>> gpurun_out/r02f_bench.err
timeout 600 python bench.py --env maze --steps 10 --warmup 3 > gpurun_out/r02f_bench_maze.json 2>> gpurun_out/r02f_bench.err
timeout 600 python bench.py --env box-delivery --obstacle-config small_columns --steps 6 --warmup 3 > gpurun_out/r02f_bench_boxdel_small_columns.json 2>> gpurun_out/r02f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02f_launches_18944.csv python bench.py --envs 18944 --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/r02f_ncu_bench.log 2>&1
for f in bench bench_ref bench_boxdel bench_areaclr bench_maze bench_boxdel_small_columns; do python -c "
import json
d=json.loads(open('gpurun_out/r02f_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'))"; done
