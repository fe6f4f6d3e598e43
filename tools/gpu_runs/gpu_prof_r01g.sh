#!/bin/bash
# ncu --set full of the final-tree step kernel at steady state, full load (18 944 envs; the 36th k_shipice_step launch:
# 33 masked phase-spreading steps + 2 warm steps are skipped).   gpurun --timeout 600 -- 'bash tools/gpu_runs/gpu_prof_r01g.sh'
mkdir -p gpurun_out
timeout 540 ncu --set full --import-source on --clock-control none -k regex:^k_shipice_step$ -s 35 -c 1 -o gpurun_out/prof_r01g_step -f python tools/bench_variant.py 18944 0.5 2 1 34 > gpurun_out/ncu_r01g.log 2>&1
tail -3 gpurun_out/ncu_r01g.log
python tools/ncu_summary.py gpurun_out/r01g_ncu_step_steady_summary.json k_shipice_step_18944_steady=gpurun_out/prof_r01g_step.ncu-rep | head -40
