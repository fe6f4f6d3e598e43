"""Timing of box-delivery-v0 steps on one GPU: env-steps/s over S steps after W warm-up steps (CUDA events on the
current stream), plus the per-env counters that explain it.  usage: boxdel_time.py N_ENVS STEPS [WARMUP]"""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from benchpush_b200.envs import BoxDeliveryVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
env = BoxDeliveryVecEnv(num_envs=n, cfg={"agent": {"action_type": "heading"}}, device=0)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); env.reset(); t1.record(); torch.cuda.synchronize()
print("reset %d envs: %.1f ms" % (n, t0.elapsed_time(t1)))
rng = np.random.default_rng(0)
acts = [torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32), device="cuda") for _ in range(steps + warm)]
for k in range(warm):
    env.step(acts[k])
torch.cuda.synchronize()
sub0 = sum(env.get_step_debug(e)["substeps"] for e in range(0, n, max(1, n // 16)))
ms = []
for k in range(steps):
    t0.record(); env.step(acts[warm + k]); t1.record(); torch.cuda.synchronize()
    ms.append(t0.elapsed_time(t1))
sub1 = sum(env.get_step_debug(e)["substeps"] for e in range(0, n, max(1, n // 16)))
print("step ms:", ["%.1f" % m for m in ms])
print("env-steps/s: %.1f   physics substeps per env-step (sampled): %.0f   env block %.1f KB" %
      (n * steps / (sum(ms) / 1e3), (sub1 - sub0) / 16 / steps, env._L.bp_boxdel_env_bytes(env._h) / 1024))
print("status max:", int(env._buf["info_i"][:, 3].max()), " boxes delivered:", int(env._buf["info_i"][:, 0].sum()))
