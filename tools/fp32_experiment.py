"""Precision experiment: what would fp32 body state / contacts / solver cost in accuracy?

The oracle (oracle/chipmunk_oracle.cpp) is built a second time with cpFloat = float: every position, velocity, contact,
impulse and solver quantity of the Chipmunk restatement in fp32, ABSOLUTE world coordinates included (the worst case: a
kernel would keep poses in fp64 and work in body-relative fp32).  Against the fp64 oracle, same layouts and actions:
  A. trajectories from reset on the reference ice fields of tests/golden: per-env-step max pose difference;
  B. the 200 seeded episodes of BASELINE config 1 on the reference's own 100 trials: success rate, efficiency, effort
     against the fp64 canonical-order golden (binomial 95 % interval, two-sample KS).
    python tools/fp32_experiment.py   ->  profiles/r02b_fp32_experiment.json
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np


def trajectories():
    from benchpush_b200.layouts import load_packed, unpack_trials
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    out = {}
    for name in ("c10_t7", "c30_t7", "c50_t0"):
        p = load_packed(os.path.join(ROOT, "tests", "golden", "ice_%s.npz" % name))
        env = ShipIceOracleEnv({"concentration": int(name[1:3]) / 100}, experiment=unpack_trials(p), order_mode=0)
        env.reset()
        rng = np.random.default_rng(1)
        rows = [env.space.state()]
        for k in range(30):
            _, _, term, _, _ = env.step(float(np.float32(rng.uniform(-1, 1))), want_obs=False)
            rows.append(env.space.state())
            if term:
                break
        out[name] = np.stack(rows)
    return out


def episodes(nproc):
    import multiprocessing as mp
    import make_golden_200ep as g
    with mp.Pool(nproc) as pool:
        rows = pool.map(g.run_episode, [(k, 0, True, 0) for k in range(g.N_EP)], chunksize=1)
    rows.sort()
    return {"steps": [r[1] for r in rows], "success": [int(r[2]) for r in rows], "efficiency": [r[3] for r in rows], "effort": [r[4] for r in rows]}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        t = trajectories()
        np.savez(sys.argv[2], **t)
        json.dump(episodes(int(sys.argv[3])), open(sys.argv[2] + ".json", "w"))
        sys.exit(0)
    nproc = os.cpu_count() or 1
    res = {}
    for tag, env in (("f64", {}), ("f32", {"BP_ORACLE_F32": "1"})):
        out = "/tmp/fp32exp_%s.npz" % tag
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "child", out, str(nproc)], env={**os.environ, **env})
        res[tag] = (np.load(out), json.load(open(out + ".json")))
    from scipy.stats import ks_2samp
    rep = {"what": __doc__.split("\n\n")[0], "trajectories": {}, "episodes": {}}
    for name in res["f64"][0].files:
        a, b = res["f64"][0][name], res["f32"][0][name]
        n = min(len(a), len(b))
        pos = np.abs(a[:n, :, :2] - b[:n, :, :2]).max(axis=(1, 2)); ang = np.abs(a[:n, :, 2] - b[:n, :, 2]).max(axis=1)
        rep["trajectories"][name] = {"steps": int(n - 1), "pos_err_m_step1": float(pos[1]), "pos_err_m_step5": float(pos[min(5, n - 1)]),
                                     "pos_err_m_step10": float(pos[min(10, n - 1)]), "pos_err_m_last": float(pos[-1]),
                                     "ang_err_rad_step1": float(ang[1]), "ang_err_rad_step10": float(ang[min(10, n - 1)]), "ang_err_rad_last": float(ang[-1])}
    e64, e32 = res["f64"][1], res["f32"][1]
    p = float(np.mean(e64["success"])); half = 1.96 * np.sqrt(p * (1 - p) / len(e64["success"]))
    rep["episodes"] = {"n": len(e64["success"]), "success_f64": p, "success_f32": float(np.mean(e32["success"])), "binomial_halfwidth": half,
                       "success_inside_interval": bool(abs(np.mean(e32["success"]) - p) <= half),
                       "steps_equal_fraction": float(np.mean(np.array(e64["steps"]) == np.array(e32["steps"]))),
                       "efficiency_mean_f64": float(np.mean(e64["efficiency"])), "efficiency_mean_f32": float(np.mean(e32["efficiency"])),
                       "effort_mean_f64": float(np.mean(e64["effort"])), "effort_mean_f32": float(np.mean(e32["effort"])),
                       "ks_p_efficiency": float(ks_2samp(e64["efficiency"], e32["efficiency"]).pvalue),
                       "ks_p_effort": float(ks_2samp(e64["effort"], e32["effort"]).pvalue),
                       "max_abs_effort_diff_per_episode": float(np.max(np.abs(np.array(e64["effort"]) - np.array(e32["effort"]))))}
    json.dump(rep, open(os.path.join(ROOT, "profiles", "r02b_fp32_experiment.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
