#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 400 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== areaclr time 16384"; timeout 600 python tools/areaclr_time.py 16384 4 2 > gpurun_out/areaclr_time_16384.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/areaclr_time_16384.log
echo "== areaclr time 4096"; timeout 600 python tools/areaclr_time.py 4096 4 2 > gpurun_out/areaclr_time_4096.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/areaclr_time_4096.log
