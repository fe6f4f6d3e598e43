"""Golden episode scores for maze-NAMO-v0 (BASELINE config 4's env): 200 seeded episodes with random actions through the CPU
oracle.  Episode k runs layout k of benchpush_b200.layouts.generate_maze_layouts(200, seed=0) (maze v1, 5 movable obstacles;
both sides regenerate it), actions a[k, t] ~ U(-1, 1) from np.random.default_rng(1) (float32, recorded), until `terminated`
(goal reached or wall hit, maze_NAMO_env.py:440-452) or 200 steps.  Two runs, as for ship-ice: libm cos / sin and the kernels'
portable routine.

    python tools/make_golden_200ep_maze.py [N_PROC]  ->  tests/golden/maze_200ep.npz
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

N_EP, MAX_STEPS, SEED_LAYOUT, SEED_ACT = 200, 200, 0, 1


def actions_table():
    return np.random.default_rng(SEED_ACT).uniform(-1, 1, size=(N_EP, MAX_STEPS)).astype(np.float32)


def layouts():
    from benchpush_b200.layouts import generate_maze_layouts
    return generate_maze_layouts(N_EP, seed=SEED_LAYOUT)


def run_episode(args):
    k, portable = args
    from oracle import cpo
    from oracle.maze_oracle import MazeNamoMetricOracle, MazeNamoOracleEnv
    cpo.set_portable_trig(bool(portable))
    acts = actions_table()[k]
    env = MazeNamoOracleEnv(layouts=[layouts()[k]], order_mode=1)
    obs, info = env.reset()
    m = MazeNamoMetricOracle("oracle", 1)
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
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maze_200ep.npz"), **out)
