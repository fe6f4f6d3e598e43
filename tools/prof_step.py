"""One reset + a few steps at a small batch, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchpush_b200.envs import ShipIceVecEnv
from benchpush_b200.layouts import generate_experiment, pack_trials
E = int(sys.argv[1]) if len(sys.argv) > 1 else 592
conc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
packed = pack_trials(generate_experiment(16, conc, seed=99))
env = ShipIceVecEnv(num_envs=E, cfg={"concentration": conc}, layouts=packed, device=0, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for k in range(steps):
    env.step_async_nocheck(torch.rand(E, generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("done")
