#!/bin/bash
# A/B: env-block accesses compiled as LDG / STG (address-space assumption in view_at, -DBP_ASSUME_GLOBAL) against generic LD / ST.
mkdir -p gpurun_out
AB=gpurun_out/r02l_assume_global_ab.txt; : > $AB
V=$PWD/benchpush_b200/csrc/variants
for n in aglobal default; do
  echo "variant $n" >> $AB
  ( [ $n != default ] && export BENCHPUSH_B200_LIB=$V/lib_$n.so; timeout 150 python tools/bench_variant.py 65536 0.5 3 4 10 ) 2>&1 | tail -1 >> $AB
done
cat $AB
