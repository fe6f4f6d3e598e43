#!/bin/bash
# compute-sanitizer memcheck over the room-env kernels (small batches) and a 2-step ship-ice run.
set -u
mkdir -p gpurun_out
echo "== memcheck box-delivery"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/boxdel_time.py 48 2 1 > gpurun_out/memcheck_boxdel.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_boxdel.log
echo "== memcheck area-clearing"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/areaclr_time.py 48 2 1 > gpurun_out/memcheck_areaclr.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/memcheck_areaclr.log
echo "== racecheck area-clearing (shared memory)"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/areaclr_time.py 8 1 1 > gpurun_out/racecheck_areaclr.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/racecheck_areaclr.log
