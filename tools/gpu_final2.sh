#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
for n in 8192 16384 65536; do
  echo "== bench envs=$n"; timeout 900 python bench.py --envs $n --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('value %.0f ms/step %.1f kernel_ms %.1f status %s'%(l['value'],l['ms_per_step'],l['roofline']['kernel_ms'],l['config']['env_status_bits']))"
done
echo "== BK equality"; timeout 600 python - <<'PY'
import os, subprocess, sys
code = """
import numpy as np, torch, hashlib
from benchpush_b200.envs import ShipIceVecEnv
from benchpush_b200.layouts import generate_experiment, pack_trials
p = pack_trials(generate_experiment(4, 0.5, seed=3))
env = ShipIceVecEnv(num_envs=64, cfg={'concentration': 0.5}, layouts=p, device=0)
env.reset()
rng = np.random.default_rng(0)
for k in range(4):
    env.step(torch.as_tensor(rng.uniform(-1, 1, size=64).astype(np.float32)))
h = hashlib.sha256()
for e in (0, 17, 63): h.update(env.get_bodies(e).tobytes())
h.update(env.obs.cpu().numpy().tobytes())
print(h.hexdigest())
"""
out = []
for bk in ("8", "4"):
    env = dict(os.environ, BP_FORCE_BK=bk)
    out.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout.strip().splitlines()[-1])
print(out, "EQUAL" if out[0] == out[1] else "DIFFERENT")
PY
