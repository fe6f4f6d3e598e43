#!/bin/bash
# Round-end verification on the GPU box: tests, smoke, bench lines (ours + reference arm), maze / observation-mode timings,
# ncu launch list of the bench command.   gpurun --timeout 1100 -- 'bash tools/gpu_final3.sh'
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_r01g.json; cut -c1-200 gpurun_out/bench_ref_r01g.json
echo "== bench"; timeout 400 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r01g.json; cut -c1-400 gpurun_out/bench_r01g.json
echo "== maze 16384"; timeout 200 python tools/maze_time.py 16384 10 2>&1 | grep -v "^env bytes" | tail -3 | tee gpurun_out/maze_time_r01g.txt
echo "== obs modes"; timeout 200 python tools/obs_modes_time.py 4096 2>&1 | tail -3 | tee gpurun_out/obs_modes_r01g.txt
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 120 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --envs 4736 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"; grep -c k_shipice gpurun_out/launches_r01g.csv
