#!/bin/bash
# Re-verification after the last edits: full GPU suite, smoke, launch list of the bench command at a batch that runs the
# main step kernel (k_shipice_step, 8 envs per warp).   gpurun --timeout 900 -- 'bash tools/gpu_final4.sh'
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 120 --csv --log-file gpurun_out/launches_r01g_18944.csv python bench.py --envs 18944 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"; grep -c k_shipice gpurun_out/launches_r01g_18944.csv
