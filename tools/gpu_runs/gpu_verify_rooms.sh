#!/bin/bash
# memcheck of the room-env paths added in round 2 + the 200-episode box-delivery tests (incl. small_columns) + full GPU suite.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_boxdel.py -m gpu -x -q -k "column" 2>&1 | tail -6 | tee gpurun_out/r02d_memcheck_rooms.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spfa.py -m gpu -x -q -k "reference_order" 2>&1 | tail -4 | tee -a gpurun_out/r02d_memcheck_rooms.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02d_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02d_gputest.log
