#!/bin/bash
# Follow-up of tools/gpu_runs/gpu_r02i.sh: which per-warp queue capacity made the small-queue variants 19 % slower?  One capacity changed
# at a time (same workload and arguments as r02i; default there: 61 724 env-steps/s), plus two larger capacities.
mkdir -p gpurun_out
AB=gpurun_out/r02j_step_queue_capacity_ab.txt; : > $AB
V=$PWD/benchpush_b200/csrc/variants
for n in bwl128 bq64 bigq bwl1024 default; do
  echo "variant $n" >> $AB
  ( [ $n != default ] && export BENCHPUSH_B200_LIB=$V/lib_$n.so; timeout 150 python tools/bench_variant.py 65536 0.5 3 4 10 ) 2>&1 | tail -1 >> $AB
done
cat $AB
