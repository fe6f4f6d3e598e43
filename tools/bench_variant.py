"""Time ship-ice steps with the CUDA library named by BENCHPUSH_B200_LIB (development aid).
usage: bench_variant.py ENVS CONC WARM STEPS -> one line: env-steps/s over STEPS steps after WARM warm-up steps,
plus a checksum of the final state (must not depend on the tuning variant)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchpush_b200.envs import ShipIceVecEnv
from benchpush_b200.layouts import generate_experiment, pack_trials
E = int(sys.argv[1]); conc = float(sys.argv[2]); warm = int(sys.argv[3]); steps = int(sys.argv[4])
desync = int(sys.argv[5]) if len(sys.argv) > 5 else 1      # > 1: spread the episode phases first (steady state, like bench.py)
packed = pack_trials(generate_experiment(32, conc, seed=99))
env = ShipIceVecEnv(num_envs=E, cfg={"concentration": conc}, layouts=packed, device=0, auto_reset=True, precision=os.environ.get("BP_PRECISION", "fp64"))
env.check_status = False
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
if desync > 1:
    phase = torch.arange(E, device="cuda") % desync
    for r in range(desync - 1):
        env.step_async_nocheck(torch.rand(E, generator=g, device="cuda") * 2 - 1, mask=(phase > r))
    torch.cuda.synchronize()
acts = [(torch.rand(E, generator=g, device="cuda") * 2 - 1) for _ in range(warm + steps)]
for k in range(warm):
    env.step_async_nocheck(acts[k])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(warm, warm + steps):
    env.step_async_nocheck(acts[k])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
h = hashlib.sha1()
for e in (0, 1, 5, E - 1):
    h.update(env.get_bodies(e).tobytes())
h.update(env.obs[:8].cpu().numpy().tobytes()); h.update(env._total_work.cpu().numpy().tobytes())
print("%s envs %d: %.1f ms/step -> %.0f env-steps/s  status %d  sha %s" % (os.path.basename(os.environ.get("BENCHPUSH_B200_LIB", "default")), E, ms / steps, E * steps / ms * 1e3, int(env._status.max()), h.hexdigest()[:12]), flush=True)
