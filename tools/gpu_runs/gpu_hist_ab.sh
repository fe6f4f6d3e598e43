#!/bin/bash
# A/B of the history-based cost key in the grouped mode of the step schedule (BP_SCHED_HISTORY=2) at the 1-GPU batch size.
mkdir -p gpurun_out; : > gpurun_out/r02h_hist65k_ab.log
for hs in 1 2; do
  echo "E=65536 BP_SCHED_HISTORY=$hs" >> gpurun_out/r02h_hist65k_ab.log
  BP_SCHED_HISTORY=$hs timeout 900 python tools/bench_variant.py 65536 0.5 3 6 34 >> gpurun_out/r02h_hist65k_ab.log 2>&1
done
cat gpurun_out/r02h_hist65k_ab.log
