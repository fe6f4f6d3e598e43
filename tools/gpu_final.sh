#!/bin/bash
# What the driver runs at round end: GPU tests, smoke, default bench (ours + reference arm).
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/bench_ref.log
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-500 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
echo "== bench area-clearing"; timeout 900 python bench.py --env area-clearing --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_areaclr.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/bench_areaclr.log
