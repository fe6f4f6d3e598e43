#!/bin/bash
# sweep: variants x blocks/SM (development aid): gpu_sweep2.sh ENVS "b1 b2 .." lib...
mkdir -p gpurun_out
E=$1; BL="$2"; shift; shift
for so in "$@"; do for b in $BL; do echo -n "blocks/SM=$b " >> gpurun_out/sweep.log; BP_BLOCKS_PER_SM=$b BENCHPUSH_B200_LIB=$PWD/$so timeout 600 python tools/bench_variant.py $E 0.5 ${WARM:-6} 3 ${DESYNC:-1} 2>&1 | tail -1 >> gpurun_out/sweep.log; done; done
cat gpurun_out/sweep.log
