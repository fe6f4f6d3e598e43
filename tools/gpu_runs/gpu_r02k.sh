#!/bin/bash
# How much of the step kernel's L1 traffic is LOCAL memory (4 288-byte stack frame: EPA hulls, out-of-line call frames)?
# A few counters of one steady-state k_shipice_step launch at full load (18 944 envs), not a full-set capture.
mkdir -p gpurun_out
M=smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,smsp__inst_executed_op_generic_ld.sum,smsp__inst_executed_op_generic_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed.sum
M=$M,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
M=$M,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 210 ncu --metrics $M --clock-control none -k regex:^k_shipice_step$ -s 11 -c 1 --csv --log-file gpurun_out/r02k_step_memory_counters.csv python tools/bench_variant.py 18944 0.5 2 1 10 > gpurun_out/r02k.log 2>&1
tail -2 gpurun_out/r02k.log
cut -d, -f13- gpurun_out/r02k_step_memory_counters.csv | tail -25
