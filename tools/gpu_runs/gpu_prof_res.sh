#!/bin/bash
# ncu --set full (+ source counters) of the resident step kernel at steady state, one full wave of envs.
mkdir -p gpurun_out
E=${1:-3552}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:^k_shipice_step_res$ -s 35 -c 1 -o gpurun_out/prof_res -f python tools/bench_variant.py $E 0.5 2 1 34 > gpurun_out/ncu_res.log 2>&1
tail -3 gpurun_out/ncu_res.log
ls -la gpurun_out/
