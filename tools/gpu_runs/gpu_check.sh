#!/bin/bash
# Full GPU check: the -m gpu suite, smoke(), one default bench run.   gpurun --timeout 1800 -- 'bash tools/gpu_runs/gpu_check.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/gputest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
