"""Golden episode scores for BASELINE config 1: ship-ice-v0, default config (concentration 0.1), 200 seeded
episodes with random actions, run through the CPU oracle.  Episode k runs trial k % 100 (ship_ice_env.py:188) of a
deterministic synthetic experiment (benchpush_b200.layouts.generate_experiment(100, 0.1, seed=99) — both sides regenerate it,
so no layout fixture is needed), with actions a[k, t] ~ U(-1, 1) from np.random.default_rng(0) (float32, recorded in the fixture),
until `terminated` or 300 steps (no TimeLimit, as evaluate() unwraps the env).

Two runs: the oracle in its default mode (libm cos / sin — what Chipmunk calls: the distributions the CUDA path has to match
statistically) and with the portable cos / sin of the kernels (per-episode scores the CUDA path has to reproduce exactly).

    python tools/make_golden_200ep.py [N_PROC]  ->  tests/golden/shipice_200ep.npz
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

N_EP, MAX_STEPS, CONC, SEED_LAYOUT, SEED_ACT = 200, 300, 0.1, 99, 0


def actions_table():
    return np.random.default_rng(SEED_ACT).uniform(-1, 1, size=(N_EP, MAX_STEPS)).astype(np.float32)


def run_episode(args):
    k, portable = args
    from benchpush_b200.layouts import generate_experiment
    from oracle import cpo
    from oracle.ship_ice_oracle import ShipIceMetricOracle, ShipIceOracleEnv
    cpo.set_portable_trig(bool(portable))
    trials = run_episode.trials if hasattr(run_episode, "trials") else generate_experiment(100, CONC, seed=SEED_LAYOUT)
    run_episode.trials = trials
    acts = actions_table()[k]
    env = ShipIceOracleEnv({"concentration": CONC}, experiment=[trials[k % 100]], order_mode=1)
    obs, info = env.reset()
    m = ShipIceMetricOracle("oracle", ship_mass=1, goal=(0, 9))
    m.reset(info)
    n = 0
    for t in range(MAX_STEPS):
        obs, r, term, trunc, info = env.step(float(acts[t]), want_obs=False)
        n += 1
        last = term or t == MAX_STEPS - 1
        m.update(info, r, last)
        if last:
            break
    return (k, n, bool(info["trial_success"]), float(m.efficiency_scores[0]), float(m.effort_scores[0]), float(m.rewards[0]),
            float(info["total_work"]))


if __name__ == "__main__":
    import multiprocessing as mp
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    out = {"actions": actions_table()}
    for portable in (0, 1):
        t0 = time.time()
        with mp.Pool(nproc) as pool:
            rows = pool.map(run_episode, [(k, portable) for k in range(N_EP)], chunksize=1)
        rows.sort()
        tag = "portable" if portable else "libm"
        out["%s_steps" % tag] = np.array([r[1] for r in rows], dtype=np.int32)
        out["%s_success" % tag] = np.array([r[2] for r in rows], dtype=np.uint8)
        out["%s_efficiency" % tag] = np.array([r[3] for r in rows])
        out["%s_effort" % tag] = np.array([r[4] for r in rows])
        out["%s_reward" % tag] = np.array([r[5] for r in rows])
        out["%s_total_work" % tag] = np.array([r[6] for r in rows])
        print(tag, "%.0f s" % (time.time() - t0), "success %.3f" % out["%s_success" % tag].mean(), "steps mean %.1f" % out["%s_steps" % tag].mean(),
              "eff %.4f effort %.4f" % (out["%s_efficiency" % tag].mean(), out["%s_effort" % tag].mean()), flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shipice_200ep.npz"), **out)
