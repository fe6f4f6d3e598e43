#!/bin/bash
# A/B of the partly-filled-batch schedule (BP_NO_PAD) at the per-GPU shares of the N = 8 / 4 strong-scaling points; box-delivery
# benches with the box-distance cache.   gpurun --timeout 2400 -- 'bash tools/gpu_runs/gpu_pad_ab.sh'
mkdir -p gpurun_out; : > gpurun_out/r02_pad_ab.log
timeout 900 python -m pytest tests/test_gpu_shipice.py tests/test_gpu_boxdel.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02_pad_ab.log
for E in 8192 16384 4096; do
  for np in 1 0; do
    echo "E=$E BP_NO_PAD=$np" >> gpurun_out/r02_pad_ab.log
    BP_NO_PAD=$np timeout 900 python tools/bench_variant.py $E 0.5 3 5 34 >> gpurun_out/r02_pad_ab.log 2>&1
  done
done
cat gpurun_out/r02_pad_ab.log
for oc in small_columns large_columns; do
  timeout 900 python bench.py --env box-delivery --obstacle-config $oc --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_boxdel_$oc.json 2> gpurun_out/r02b_bench_boxdel_$oc.err
  tail -n 1 gpurun_out/r02b_bench_boxdel_$oc.json | cut -c1-230; tail -2 gpurun_out/r02b_bench_boxdel_$oc.err
done
