#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/fp32_bk.log
for bk in 12 16; do
  echo "fp32 batch BK=$bk" >> gpurun_out/fp32_bk.log
  BENCHPUSH_B200_LIB_F32=benchpush_b200/csrc/variants/f32_bk$bk.so BP_PRECISION=fp32 BP_STEP_IMPL=batch timeout 900 python tools/bench_variant.py 65536 0.5 3 5 34 >> gpurun_out/fp32_bk.log 2>&1
done
cat gpurun_out/fp32_bk.log
