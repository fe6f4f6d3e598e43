#!/bin/bash
# A/B of library variants at steady state (development aid):
#   gpurun -- 'bash tools/gpu_ab2.sh ENVS lib1.so lib2.so ...'   ("default" = the in-tree library)
mkdir -p gpurun_out; : > gpurun_out/ab2.log
E=${1:-18944}; shift
if [ "${RUN_TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset BENCHPUSH_B200_LIB; else export BENCHPUSH_B200_LIB=$PWD/$lib; fi
  timeout 600 python tools/bench_variant.py $E 0.5 2 3 34 >> gpurun_out/ab2.log 2>&1
done
cat gpurun_out/ab2.log
