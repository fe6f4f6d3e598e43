#!/bin/bash
# Address-space assumptions in the batched step (BP_ASSUME_GLOBAL_PTR / layout_g): throughput of the two builds at the r02i / r02l
# workload (generic-access build there: 61 361 - 61 724 env-steps/s), then the ship-ice / maze / room-env GPU tests on the faster one.
mkdir -p gpurun_out
AB=gpurun_out/r02m_address_space_ab.txt; : > $AB
V=$PWD/benchpush_b200/csrc/variants
for n in default nolayoutg; do
  echo "variant $n" >> $AB
  ( [ $n != default ] && export BENCHPUSH_B200_LIB=$V/lib_$n.so; timeout 100 python tools/bench_variant.py 65536 0.5 3 4 10 ) 2>&1 | tail -1 >> $AB
done
W=$(python - <<'P'
import re
r={}; lab=None
for ln in open("gpurun_out/r02m_address_space_ab.txt"):
    if ln.startswith("variant "): lab=ln.split()[1]; continue
    m=re.search(r"-> (\d+) env-steps/s", ln)
    if m and lab: r[lab]=int(m.group(1))
print("nolayoutg" if r.get("nolayoutg",0) > r.get("default",0) else "default")
P
)
echo "tests run with: $W" >> $AB
[ "$W" = nolayoutg ] && export BENCHPUSH_B200_LIB=$V/lib_nolayoutg.so
python -c "from oracle import cpo; cpo.lib()" 2>&1 | tail -1
timeout 150 python -m pytest tests/test_gpu_shipice.py tests/test_gpu_maze.py tests/test_gpu_boxdel.py tests/test_gpu_low_dim.py tests/test_gpu_areaclr.py -m gpu -q -n 6 2>&1 | tail -4 >> $AB
cat $AB
