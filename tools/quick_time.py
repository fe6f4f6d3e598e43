"""Quick device timing of ship-ice steps (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchpush_b200.envs import ShipIceVecEnv
from benchpush_b200.layouts import generate_experiment, pack_trials

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
conc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t = time.time(); packed = pack_trials(generate_experiment(16, conc, seed=99)); print("layouts %.1fs" % (time.time() - t), flush=True)
env = ShipIceVecEnv(num_envs=E, cfg={"concentration": conc}, layouts=packed, device=0, auto_reset=True)
print("env bytes", env.env_bytes(), "total GB", env.env_bytes() * E / 1e9, flush=True)
torch.cuda.synchronize(); t = time.time(); env.reset(); torch.cuda.synchronize(); print("reset %.3fs" % (time.time() - t), flush=True)
g = torch.Generator(device="cuda"); g.manual_seed(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(steps):
    a = torch.rand(E, generator=g, device="cuda") * 2 - 1
    ev0.record(); env.step_async_nocheck(a); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print("step %d: %.2f ms  -> %.0f env-steps/s   status %d" % (k, ms, E / ms * 1e3, int(env._status.max())), flush=True)
tot = np.zeros(8); cs = []
for e in range(0, E, max(1, E // 64)):
    c = env.counters(e); cs.append(c); tot += np.array(c["phase_clocks"], dtype=np.float64)
if tot.sum() > 0:
    names = ["integrate", "rebuild", "eval_pairs", "wake+order", "prestep+damp", "solve", "-", "-"]
    print("phase share:", {n: round(float(t / tot.sum()), 3) for n, t in zip(names, tot) if n != "-"}, "clocks/substep %.0f" % (tot.sum() / sum(c["substeps"] for c in cs)))
c = cs[0]; print(c)
sub = max(c["substeps"], 1)
print("per substep: awake %.1f live %.1f pairs_eval %.1f gjk %.1f recolour %.3f" % (c["awake_sum"] / sub, c["live_sum"] / sub, c["pairs_eval"] / sub, c["gjk"] / sub, c["recolour"] / sub))
