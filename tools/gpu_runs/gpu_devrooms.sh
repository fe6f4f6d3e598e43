#!/bin/bash
# Device-built rooms (csrc/boxdel_room.cuh) on one B200: parity tests, memcheck of the room build, reset timing, bench lines.
mkdir -p gpurun_out; : > gpurun_out/r02e_devrooms.log
timeout 900 python -m pytest tests/test_gpu_boxdel.py tests/test_gpu_areaclr.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/r02e_devrooms.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_boxdel.py -m gpu -x -q -k "column and small" 2>&1 | tail -4 >> gpurun_out/r02e_devrooms.log
timeout 600 python - >> gpurun_out/r02e_devrooms.log 2>&1 <<'PY'
import time, torch, numpy as np
from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
for oc in ("small_columns", "large_divider"):
    cfg = {"agent": {"action_type": "heading"}, "env": {"obstacle_config": oc}}
    t = time.time(); env = BoxDeliveryVecEnv(num_envs=4096, cfg=cfg, device=0); t_build = time.time() - t
    torch.cuda.synchronize(); t = time.time(); env.reset(); torch.cuda.synchronize(); t1 = time.time() - t
    t = time.time(); env.reset(); torch.cuda.synchronize(); t2 = time.time() - t
    print("%s: 4096 envs, %d layouts; host setup %.1f s; reset incl. room build %.3f s, second reset %.3f s; env bytes %d" % (oc, len(env.layouts), t_build, t1, t2, env._L.bp_boxdel_env_bytes(env._h)))
    env.close()
PY
for args in "--env box-delivery --obstacle-config small_columns" "--env box-delivery --obstacle-config small_columns --host-rooms" \
            "--env box-delivery --obstacle-config large_columns" "--env box-delivery --obstacle-config large_divider"; do
  echo "bench.py $args" >> gpurun_out/r02e_devrooms.log
  tag=$(echo "$args" | tr -d ' -' )
  timeout 900 python bench.py $args --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02e_err.txt | tail -n 1 > gpurun_out/r02e_bench_$tag.json
  tail -2 gpurun_out/r02e_err.txt >> gpurun_out/r02e_devrooms.log
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02e_bench_$tag.json').read())
print(round(d['value']), 'env-steps/s  ms/step', round(d['ms_per_step'],1), 'move', round(d['roofline']['kernel_ms'],1), 'obs', round(d['roofline']['obs_kernel_ms'],1), 'e2e', round(d['e2e']['value']), 'status', d['config']['env_status_bits'])" >> gpurun_out/r02e_devrooms.log 2>&1
done
cat gpurun_out/r02e_devrooms.log
