#!/bin/bash
# occupancy sweep + full-load profile of the step kernel (development aid)
mkdir -p gpurun_out; : > gpurun_out/occ.log
for b in 1 2 3 4; do echo "blocks/SM=$b" >> gpurun_out/occ.log; BP_BLOCKS_PER_SM=$b timeout 600 python tools/bench_variant.py ${1:-16384} 0.5 6 3 >> gpurun_out/occ.log 2>&1; done
cat gpurun_out/occ.log
timeout 900 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section InstructionStats --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_shipice_step -s 5 -c 1 -o gpurun_out/prof_step_full -f python tools/bench_variant.py 4736 0.5 6 1 > gpurun_out/ncu_full_load.log 2>&1; tail -3 gpurun_out/ncu_full_load.log
