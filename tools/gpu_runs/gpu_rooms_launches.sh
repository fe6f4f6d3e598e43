#!/bin/bash
# ncu launch list of box-delivery small_columns (device-built rooms, reference-order parents): reset + steps at 4096 envs;
# plus the GPU tests of the room envs and smoke on the final tree.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py tests/test_gpu_spfa.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02g_rooms_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r02g_rooms_test.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02g_boxdel_small_columns_launches.csv python bench.py --env box-delivery --obstacle-config small_columns --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02g_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02g_boxdel_small_columns_launches.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv: agg[r[kn].split('(')[0]].append(float(r[mv].replace(',', '')) / 1e6)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print(k, len(v), 'launches', ' '.join('%.1f' % x for x in v[:10]), 'ms')
PY
