#!/bin/bash
# Full GPU check: the -m gpu suite, smoke(), one default bench run.   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee -a gpurun_out/gputest.log
timeout 900 python bench.py --steps ${1:-6} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
