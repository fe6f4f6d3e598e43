"""Golden episode scores for BASELINE config 1: ship-ice-v0, default config (concentration 0.1), 200 seeded
episodes with random actions, run through the CPU oracle.  Episode k runs trial k % 100 (ship_ice_env.py:188) of a
deterministic synthetic experiment (benchpush_b200.layouts.generate_experiment(100, 0.1, seed=99) — both sides regenerate it,
so no layout fixture is needed), with actions a[k, t] ~ U(-1, 1) from np.random.default_rng(0) (float32, recorded in the fixture),
until `terminated` or 300 steps (no TimeLimit, as evaluate() unwraps the env).

Two runs: the oracle in its default mode (libm cos / sin — what Chipmunk calls: the distributions the CUDA path has to match
statistically) and with the portable cos / sin of the kernels (per-episode scores the CUDA path has to reproduce exactly).

    python tools/make_golden_200ep.py [N_PROC]  ->  tests/golden/shipice_200ep.npz

`--reference-data`: the same protocol on the REFERENCE'S OWN ice fields — the 100 trials of
ship_ice_nav/ice_environments/experiments_10_100_r06_d40x12.pk in the reference's trial order (episode k on trial k % 100,
ship_ice_env.py:188), packed once into tests/golden/ice_c10_all100.npz (needs /root/reference the first time) — and a third run
with the oracle's CANONICAL solver order (order_mode 0: arbiters in ascending shape pair, libm cos / sin), the order-independent
statement of the reference; the kernels solve in greedy-colour order, so that run is what quantifies the ordering error at
episode level.   ->  tests/golden/shipice_200ep_refdata.npz
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

N_EP, MAX_STEPS, CONC, SEED_LAYOUT, SEED_ACT = 200, 300, 0.1, 99, 0


def actions_table():
    return np.random.default_rng(SEED_ACT).uniform(-1, 1, size=(N_EP, MAX_STEPS)).astype(np.float32)


GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_PACK = os.path.join(GOLD, "ice_c10_all100.npz")


def reference_trials():
    """The 100 trials of the reference's concentration-0.1 experiment, from the committed pack (made from the pickle once)."""
    from benchpush_b200.layouts import load_packed, pack_trials, save_packed, unpack_trials
    if not os.path.exists(REF_PACK):
        import ref_data
        exp = ref_data.load_experiment(CONC)
        save_packed(REF_PACK, pack_trials([exp[i] for i in range(100)]))
    return unpack_trials(load_packed(REF_PACK))


def run_episode(args):
    k, portable = args[:2]
    refdata = len(args) > 2 and args[2]
    order_mode = args[3] if len(args) > 3 else 1
    from benchpush_b200.layouts import generate_experiment
    from oracle import cpo
    from oracle.ship_ice_oracle import ShipIceMetricOracle, ShipIceOracleEnv
    cpo.set_portable_trig(bool(portable))
    key = "trials_ref" if refdata else "trials"
    if not hasattr(run_episode, key):
        setattr(run_episode, key, reference_trials() if refdata else generate_experiment(100, CONC, seed=SEED_LAYOUT))
    trials = getattr(run_episode, key)
    acts = actions_table()[k]
    env = ShipIceOracleEnv({"concentration": CONC}, experiment=[trials[k % 100]], order_mode=order_mode)
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
    refdata = "--reference-data" in sys.argv
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    nproc = int(argv[0]) if argv else os.cpu_count()
    if refdata:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        reference_trials()
    out = {"actions": actions_table()}
    for portable, order_mode in ((0, 1), (1, 1)) + (((0, 0),) if refdata else ()):
        t0 = time.time()
        with mp.Pool(nproc) as pool:
            rows = pool.map(run_episode, [(k, portable, refdata, order_mode) for k in range(N_EP)], chunksize=1)
        rows.sort()
        tag = "canonical" if order_mode == 0 else ("portable" if portable else "libm")
        out["%s_steps" % tag] = np.array([r[1] for r in rows], dtype=np.int32)
        out["%s_success" % tag] = np.array([r[2] for r in rows], dtype=np.uint8)
        out["%s_efficiency" % tag] = np.array([r[3] for r in rows])
        out["%s_effort" % tag] = np.array([r[4] for r in rows])
        out["%s_reward" % tag] = np.array([r[5] for r in rows])
        out["%s_total_work" % tag] = np.array([r[6] for r in rows])
        print(tag, "%.0f s" % (time.time() - t0), "success %.3f" % out["%s_success" % tag].mean(), "steps mean %.1f" % out["%s_steps" % tag].mean(),
              "eff %.4f effort %.4f" % (out["%s_efficiency" % tag].mean(), out["%s_effort" % tag].mean()), flush=True)
    np.savez_compressed(os.path.join(GOLD, "shipice_200ep_refdata.npz" if refdata else "shipice_200ep.npz"), **out)
