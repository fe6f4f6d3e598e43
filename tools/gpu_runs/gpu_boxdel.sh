#!/bin/bash
# One GPU-box visit for box-delivery-v0: parity tests, timing at BASELINE config 2 (4096 envs), ncu launch list and
# full captures of the two heaviest kernels.   gpurun --timeout 1500 -- 'bash tools/gpu_runs/gpu_boxdel.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest boxdel"; timeout 900 python -m pytest tests/test_gpu_boxdel.py -x -q > gpurun_out/pytest_boxdel.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_boxdel.log
echo "== time 4096"; timeout 600 python tools/boxdel_time.py 4096 3 1 > gpurun_out/boxdel_time_4096.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/boxdel_time_4096.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/boxdel_launches.csv python tools/boxdel_time.py 1024 2 1 > gpurun_out/boxdel_ncu_launch.log 2>&1; echo "rc=$?"; grep -c k_bd gpurun_out/boxdel_launches.csv
if [ "${NCU_FULL:-1}" = "1" ]; then
echo "== ncu full k_bd_dist"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bd_dist -s 2 -c 1 -o gpurun_out/prof_bd_dist -f python tools/boxdel_time.py 1024 2 1 > gpurun_out/ncu_bd_dist.log 2>&1; echo "rc=$?"
echo "== ncu full k_bd_move"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bd_move -s 1 -c 1 -o gpurun_out/prof_bd_move -f python tools/boxdel_time.py 1024 2 1 > gpurun_out/ncu_bd_move.log 2>&1; echo "rc=$?"
echo "== ncu full k_bd_obs"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bd_obs -s 1 -c 1 -o gpurun_out/prof_bd_obs -f python tools/boxdel_time.py 1024 2 1 > gpurun_out/ncu_bd_obs.log 2>&1; echo "rc=$?"
fi
ls -la gpurun_out | head -40
