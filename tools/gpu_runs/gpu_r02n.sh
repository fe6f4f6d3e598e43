#!/bin/bash
# Confirmation of the final default build (env-block accesses as LDG / STG): r02i workload, then the ship-ice GPU tests.
mkdir -p gpurun_out
O=gpurun_out/r02n_final_build_check.txt
timeout 60 python tools/bench_variant.py 65536 0.5 3 4 10 2>&1 | tail -1 > $O
timeout 45 python -m pytest tests/test_gpu_shipice.py tests/test_gpu_maze.py -m gpu -q -n 8 2>&1 | tail -2 >> $O
cat $O
