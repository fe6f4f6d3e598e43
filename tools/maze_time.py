"""Device timing of maze-NAMO steps (development aid): ENVS STEPS -> env-steps/s, kernel split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchpush_b200.envs import MazeNamoVecEnv
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
env = MazeNamoVecEnv(num_envs=E, device=0, auto_reset=True, n_synthetic_layouts=256)
env.check_status = False
env.reset()
env.set_profiling(True)
g = torch.Generator(device="cuda"); g.manual_seed(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    acts = [(torch.rand(E, generator=g, device="cuda") * 2 - 1) for _ in range(steps)]
    torch.cuda.synchronize(); env.kernel_ms()
    e0.record()
    for a in acts:
        env.step_async_nocheck(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); km = env.kernel_ms()
    print("rep %d: %.2f ms/step -> %.0f env-steps/s  (step kernel %.2f ms, obs kernel %.2f ms per launch)  status %d  terminated %.3f" % (
        rep, ms / steps, E * steps / ms * 1e3, km["step_ms"] / max(km["step_launches"], 1), km["obs_ms"] / max(km["obs_launches"], 1),
        int(env._status.max()), float(env._terminated.float().mean())), flush=True)
print("env bytes", env.env_bytes(), "counters env0", env.counters(0))
