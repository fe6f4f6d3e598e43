#!/bin/bash
# Steady-state full-load source-level profile of the step kernel (development aid).
#   gpurun --timeout 1500 -- 'bash tools/gpu_prof_steady.sh [ENVS] [DESYNC]'
E=${1:-18944}; D=${2:-34}
mkdir -p gpurun_out
# launches before the profiled one: reset (1) + desync-1 masked steps + 2 warm steps; profile the next step
S=$((D - 1 + 2))
timeout 1400 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section InstructionStats --section SourceCounters --import-source on --metrics sm__icc_request_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_shipice_step -s $S -c 1 -o gpurun_out/prof_steady -f python tools/bench_variant.py $E 0.5 2 1 $D > gpurun_out/ncu_steady.log 2>&1
tail -3 gpurun_out/ncu_steady.log
ncu -i gpurun_out/prof_steady.ncu-rep --page source --csv > gpurun_out/prof_steady_src.csv 2>/dev/null
ncu -i gpurun_out/prof_steady.ncu-rep --page raw --csv > gpurun_out/prof_steady_raw.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/prof_steady_src.csv benchpush_b200/csrc/libbenchpush_b200.so k_shipice_step8 60 > gpurun_out/prof_steady_lines.txt 2>&1
head -30 gpurun_out/prof_steady_lines.txt
