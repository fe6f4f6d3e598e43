#!/bin/bash
# Whole GPU suite + smoke + box-delivery timing / launch list.   gpurun --timeout 1800 -- 'bash tools/gpu_round2.sh'
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 400 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== boxdel time 4096"; timeout 600 python tools/boxdel_time.py 4096 4 2 > gpurun_out/boxdel_time_4096.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/boxdel_time_4096.log
echo "== boxdel time 16384"; timeout 600 python tools/boxdel_time.py 16384 3 1 > gpurun_out/boxdel_time_16384.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/boxdel_time_16384.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/boxdel_launches.csv python tools/boxdel_time.py 4096 2 1 > gpurun_out/boxdel_ncu_launch.log 2>&1; echo "rc=$?"; grep "k_bd" gpurun_out/boxdel_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
