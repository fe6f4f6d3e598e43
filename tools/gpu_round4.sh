#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest boxdel"; timeout 900 python -m pytest tests/test_gpu_boxdel.py -x -q > gpurun_out/pytest_boxdel.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_boxdel.log
echo "== boxdel time 4096"; timeout 600 python tools/boxdel_time.py 4096 8 3 > gpurun_out/boxdel_time_4096.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/boxdel_time_4096.log
echo "== bench box-delivery"; timeout 900 python bench.py --env box-delivery --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_boxdel.log 2> gpurun_out/bench_boxdel.err; echo "rc=$?"; cut -c1-1200 gpurun_out/bench_boxdel.log; tail -3 gpurun_out/bench_boxdel.err
