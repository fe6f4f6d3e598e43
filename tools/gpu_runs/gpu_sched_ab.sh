#!/bin/bash
# A/B of the step schedule (k_schedule) at the 1-GPU and the 8-GPU-share batch sizes; state checksums must agree.
mkdir -p gpurun_out; : > gpurun_out/sched_ab.log
timeout 600 python -m pytest tests/test_gpu_shipice.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/sched_ab.log
for E in 8192 65536; do
  for ns in 1 0; do
    echo "E=$E BP_NO_SCHEDULE=$ns" >> gpurun_out/sched_ab.log
    BP_NO_SCHEDULE=$ns timeout 900 python tools/bench_variant.py $E 0.5 3 5 34 >> gpurun_out/sched_ab.log 2>&1
  done
done
cat gpurun_out/sched_ab.log
