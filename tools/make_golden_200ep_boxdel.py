"""Golden episode scores for box-delivery-v0 (BASELINE config 2's env): 200 seeded episodes with random heading actions through
the CPU oracle, scored with TaskDrivenMetric.  Episode k runs sample_layout(np.random.RandomState(42 + k)) (small_empty, 10
boxes — the reference's rejection sampler and seed convention, box_delivery_env.py:313-320), actions a[k, t] ~ U(-1, 1) from
np.random.default_rng(3) (float32, recorded), until terminated / truncated (inactivity cutoff) or 100 steps (the oracle runs ~5 env steps/s per core).

    python tools/make_golden_200ep_boxdel.py [N_PROC]  ->  tests/golden/boxdel_200ep.npz
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

N_EP, MAX_STEPS, SEED_LAYOUT, SEED_ACT = 200, 100, 42, 3
# BOXDEL_OC=small_columns (or large_columns / large_divider): the same protocol in a room with random obstacles ->
# tests/golden/boxdel_200ep_<config>.npz (the per-box spfa solves make the oracle ~3x slower there)
OC = os.environ.get("BOXDEL_OC", "small_empty")
CFG = {"agent": {"action_type": "heading"}} if OC == "small_empty" else {"agent": {"action_type": "heading"}, "env": {"obstacle_config": OC}}


def actions_table():
    return np.random.default_rng(SEED_ACT).uniform(-1, 1, size=(N_EP, MAX_STEPS)).astype(np.float32)


def layout(k):
    from oracle.box_delivery_oracle import sample_layout
    return sample_layout(np.random.RandomState(SEED_LAYOUT + k), CFG)


def batched(info, reset):
    """Oracle info (one env) -> the batched layout TaskDrivenMetric consumes (leading dimension 1)."""
    polys = [np.asarray(p, dtype=np.float64) for p in info["obs"]]
    obs = np.asarray(polys)[None] if polys else np.zeros((1, 0, 4, 2))
    out = {"state": np.asarray([info["state"]], dtype=np.float64), "total_work": np.asarray([info["total_work"]], dtype=np.float64),
           "obs": obs, "n_boxes_left": np.asarray([obs.shape[1]]), "goal_positions": info["goal_positions"],
           "box_completed_statuses": np.asarray([info["box_completed_statuses"]], dtype=bool)}
    return out


def run_episode(k):
    from benchpush_b200.metrics import TaskDrivenMetric
    from oracle.box_delivery_oracle import BoxDeliveryOracleEnv
    acts = actions_table()[k]
    env = BoxDeliveryOracleEnv(cfg=CFG, layouts=[layout(k)])
    obs, info = env.reset()
    m = TaskDrivenMetric("oracle", robot_mass=1)
    m.reset(batched(info, True))
    n = 0
    for t in range(MAX_STEPS):
        obs, r, term, trunc, info = env.step(float(acts[t]), want_obs=False)
        n += 1
        last = bool(term or trunc or t == MAX_STEPS - 1)
        m.update(batched(info, False), np.asarray([r]), np.asarray([last]))
        if last:
            break
    return (k, n, float(m.success_rates[0]), float(m.efficiency_scores[0]), float(m.effort_scores[0]), float(m.rewards[0]), float(info["total_work"]))


if __name__ == "__main__":
    import multiprocessing as mp
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    t0 = time.time()
    with mp.Pool(nproc) as pool:
        rows = pool.map(run_episode, list(range(N_EP)), chunksize=1)
    rows.sort()
    out = {"actions": actions_table(), "steps": np.array([r[1] for r in rows], dtype=np.int32), "success": np.array([r[2] for r in rows]),
           "efficiency": np.array([r[3] for r in rows]), "effort": np.array([r[4] for r in rows]), "reward": np.array([r[5] for r in rows]),
           "total_work": np.array([r[6] for r in rows])}
    print("%.0f s" % (time.time() - t0), "success %.3f" % out["success"].mean(), "steps mean %.1f" % out["steps"].mean(),
          "eff %.4f effort %.4f reward %.3f" % (out["efficiency"].mean(), out["effort"].mean(), out["reward"].mean()), flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "boxdel_200ep.npz" if OC == "small_empty" else "boxdel_200ep_%s.npz" % OC), **out)
