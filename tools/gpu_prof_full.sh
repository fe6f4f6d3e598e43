#!/bin/bash
# full-load profile of the step kernel: E envs, skip S launches (development aid)
E=${1:-18944}; S=${2:-5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shipice.py -m gpu -x -q -k heterogeneous 2>&1 | grep -E "Error|assert|Mismatch|Max abs|tests/" | head -20
timeout 1500 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section InstructionStats --section SourceCounters --import-source on --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:k_shipice_step -s $S -c 1 -o gpurun_out/prof_step_full -f python tools/bench_variant.py $E 0.5 $S 1 > gpurun_out/ncu_full_load.log 2>&1; tail -3 gpurun_out/ncu_full_load.log
