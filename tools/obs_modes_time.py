"""Device timing of the ship-ice observation modes (development aid): ENVS -> ms per launch of the egocentric, global and
low-dimensional observation kernels, with the bytes each writes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchpush_b200.envs import ShipIceVecEnv
from benchpush_b200.layouts import generate_experiment, pack_trials
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
packed = pack_trials(generate_experiment(16, 0.5, seed=99))
for mode, cfg in (("egocentric", {}), ("global", {"egocentric_obs": False}), ("low_dim", {"low_dim_state": True, "fixed_trial_idx": 3})):
    env = ShipIceVecEnv(num_envs=E, cfg=dict(cfg, concentration=0.5), layouts=packed, device=0)
    env.check_status = False
    env.reset()
    a = torch.zeros(E, device="cuda")
    for k in range(2):
        env.step(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    torch.cuda.synchronize()
    if mode == "egocentric":
        env.set_profiling(True); env.kernel_ms()
        for k in range(reps):
            env.step(a)
        km = env.kernel_ms()
        ms = km["obs_ms"] / max(km["obs_launches"], 1)
        nbytes = env.obs[0].numel()
    else:
        e0.record()
        for k in range(reps):
            env._observation()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = (env.obs[0].numel() if mode == "global" else env.low_dim_obs[0].numel() * 8)
    print("%-10s envs %d: %.3f ms per launch = %.3f us/env, %d B/env written -> %.1f GB/s" % (mode, E, ms, ms * 1e3 / E, nbytes, nbytes * E / ms / 1e6), flush=True)
    env.close()
