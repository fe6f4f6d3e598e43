#!/bin/bash
# A/B of the history-based cost key of the balanced step schedule (BP_SCHED_HISTORY=1); state hashes must agree.
mkdir -p gpurun_out; : > gpurun_out/r02h_hist_ab.log
for E in 8192 4096 16384; do
  for hs in 0 1; do
    echo "E=$E BP_SCHED_HISTORY=$hs" >> gpurun_out/r02h_hist_ab.log
    BP_SCHED_HISTORY=$hs timeout 900 python tools/bench_variant.py $E 0.5 3 6 34 >> gpurun_out/r02h_hist_ab.log 2>&1
  done
done
cat gpurun_out/r02h_hist_ab.log
