#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench box-delivery"; timeout 900 python bench.py --env box-delivery --steps 4 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_boxdel.log 2> gpurun_out/bench_boxdel.err; echo "rc=$?"; cat gpurun_out/bench_boxdel.log | cut -c1-1500; tail -3 gpurun_out/bench_boxdel.err
echo "== bench box-delivery reference"; timeout 600 python bench.py --env box-delivery --impl reference --steps 3 --warmup 1 > gpurun_out/bench_boxdel_ref.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/bench_boxdel_ref.log
echo "== ncu full k_bd_move 4096"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bd_move -s 1 -c 1 -o gpurun_out/prof_bd_move_4096 -f python tools/boxdel_time.py 4096 2 1 > gpurun_out/ncu_bd_move.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bd_move.log
ls -la gpurun_out | head -30
