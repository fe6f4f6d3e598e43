#!/bin/bash
# Round-end measurements: the three bench lines (ours + reference arms), launch lists and ncu captures of the room-env kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== bench ship-ice"; timeout 1500 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
echo "== bench ship-ice reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/bench_ref.log
echo "== bench box-delivery"; timeout 900 python bench.py --env box-delivery --steps 6 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_boxdel.log 2> gpurun_out/bench_boxdel.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_boxdel.log
echo "== bench box-delivery 16384"; timeout 900 python bench.py --env box-delivery --envs 16384 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_boxdel_16384.log 2> gpurun_out/bench_boxdel16.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_boxdel_16384.log
echo "== bench area-clearing"; timeout 900 python bench.py --env area-clearing --steps 4 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_areaclr.log 2> gpurun_out/bench_areaclr.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_areaclr.log
echo "== bench area-clearing reference"; timeout 600 python bench.py --env area-clearing --impl reference --steps 3 --warmup 1 > gpurun_out/bench_areaclr_ref.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/bench_areaclr_ref.log
echo "== bench box-delivery reference"; timeout 600 python bench.py --env box-delivery --impl reference --steps 3 --warmup 1 > gpurun_out/bench_boxdel_ref.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/bench_boxdel_ref.log
echo "== ncu launch list areaclr"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/areaclr_launches.csv python tools/areaclr_time.py 4096 2 1 > gpurun_out/areaclr_ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu launch list boxdel"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/boxdel_launches.csv python tools/boxdel_time.py 4096 2 1 > gpurun_out/boxdel_ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full k_bd_move (area-clearing 4096)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bd_move -s 1 -c 1 -o gpurun_out/prof_ac_move_4096 -f python tools/areaclr_time.py 4096 2 1 > gpurun_out/ncu_ac_move.log 2>&1; echo "rc=$?"
echo "== ncu full k_bd_obs (area-clearing 4096)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bd_obs -s 1 -c 1 -o gpurun_out/prof_ac_obs_4096 -f python tools/areaclr_time.py 4096 2 1 > gpurun_out/ncu_ac_obs.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
