#!/bin/bash
# One GPU-box visit: parity tests, smoke, quick timing, bench, ncu launch list.  Run under gpurun:
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== quick_time 4096"; timeout 600 python tools/quick_time.py 4096 0.5 6 > gpurun_out/quick_4096.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/quick_4096.log
echo "== bench"; timeout 1200 python bench.py --steps ${BENCH_STEPS:-4} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py 2368 0.5 3 > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/launches.csv
if [ "${NCU_FULL:-1}" = "1" ]; then
echo "== ncu full (step kernel)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shipice_step -s 1 -c 1 -o gpurun_out/prof_step -f python tools/prof_step.py 4736 0.5 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
echo "== ncu full (obs kernel)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shipice_obs -s 2 -c 1 -o gpurun_out/prof_obs -f python tools/prof_step.py 2368 0.5 3 > gpurun_out/ncu_full_obs.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full_obs.log
fi
ls -la gpurun_out
