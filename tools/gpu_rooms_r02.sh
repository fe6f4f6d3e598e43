#!/bin/bash
# box-delivery obstacle rooms on one B200: spfa exact-parents tests + bench lines (deterministic vs reference-order parents).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spfa.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_spfa_test.log
for oc in small_columns large_columns large_divider; do
  timeout 900 python bench.py --env box-delivery --obstacle-config $oc --steps 4 --warmup 3 --cpu-seconds 8 > gpurun_out/r02_bench_boxdel_$oc.json 2> gpurun_out/r02_bench_boxdel_$oc.err
  tail -n 1 gpurun_out/r02_bench_boxdel_$oc.json | cut -c1-330; tail -2 gpurun_out/r02_bench_boxdel_$oc.err
done
timeout 900 python bench.py --env box-delivery --obstacle-config small_columns --exact-parents --envs 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_boxdel_small_columns_exact.json 2> gpurun_out/r02_bench_boxdel_small_columns_exact.err
tail -n 1 gpurun_out/r02_bench_boxdel_small_columns_exact.json | cut -c1-330; tail -2 gpurun_out/r02_bench_boxdel_small_columns_exact.err
timeout 600 python bench.py --env box-delivery --obstacle-config small_columns --envs 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_boxdel_small_columns_1024.json 2>&1
tail -n 1 gpurun_out/r02_bench_boxdel_small_columns_1024.json | cut -c1-330
