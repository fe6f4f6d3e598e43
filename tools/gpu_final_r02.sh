#!/bin/bash
# Round-2 measurement set on one B200: GPU test suite, smoke, headline bench, secondary env benches, ncu launch list and one
# full capture of the step kernel at steady state.   gpurun --timeout 3000 -- 'bash tools/gpu_final_r02.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02_gputest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -1 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench.err
timeout 600 python bench.py --env box-delivery --steps 6 --warmup 3 > gpurun_out/r02_bench_boxdel.json 2>> gpurun_out/r02_bench.err
timeout 600 python bench.py --env area-clearing --steps 6 --warmup 3 > gpurun_out/r02_bench_areaclr.json 2>> gpurun_out/r02_bench.err
timeout 300 python tools/maze_time.py 16384 > gpurun_out/r02_maze_16384.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02_launches_18944.csv python bench.py --envs 18944 --steps 2 --warmup 3 --no-cpu-baseline --no-fp32 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:^k_shipice_step$ -s 35 -c 1 -o gpurun_out/r02_prof_step -f python tools/bench_variant.py 18944 0.5 2 1 34 > gpurun_out/r02_ncu_step.log 2>&1
tail -2 gpurun_out/r02_ncu_step.log
cat gpurun_out/r02_bench.json
