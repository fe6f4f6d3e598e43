"""Timing of area-clearing-v0 steps on one GPU (BASELINE config 3: 16 384 envs).  usage: areaclr_time.py N_ENVS STEPS [WARMUP]"""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from benchpush_b200.envs import AreaClearingVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
env = AreaClearingVecEnv(num_envs=n, device=0)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); env.reset(); t1.record(); torch.cuda.synchronize()
print("reset %d envs: %.1f ms" % (n, t0.elapsed_time(t1)))
rng = np.random.default_rng(0)
acts = [torch.as_tensor(rng.uniform(-1, 1, size=n).astype(np.float32), device="cuda") for _ in range(steps + warm)]
for k in range(warm):
    env.step(acts[k])
torch.cuda.synchronize()
env.kernel_ms(); env.set_profiling(True)
ms = []
for k in range(steps):
    t0.record(); env.step(acts[warm + k]); t1.record(); torch.cuda.synchronize()
    ms.append(t0.elapsed_time(t1))
km = env.kernel_ms()
print("step ms:", ["%.1f" % m for m in ms], " k_bd_move ms/step %.1f  k_bd_obs ms/step %.1f" % (km["move_ms"] / max(km["steps"], 1), km["obs_ms"] / max(km["steps"], 1)))
print("env-steps/s: %.1f   boxes cleared: %d   status bits: %d" % (n * steps / (sum(ms) / 1e3), int(env._buf["info_i"][:, 0].sum()), int(env._buf["info_i"][:, 3].max())))
