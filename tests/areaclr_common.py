"""Shared helpers of the area-clearing tests (CPU emulation and GPU)."""
import numpy as np

CLEARING_LAYOUT = {"start": (0.0, 1.5, np.pi / 2), "boxes": [[0.0, 4.2], [-3.0, -3.0], [3.0, -2.0]]}
CLEARING_CFG = {"num_obstacles": 3}


def compare_step(env, ora, e, action, step_out, state_tol=1e-7, max_obs_diff=4, allow_path_ties=False):
    """allow_path_ties: spfa parents are queue-order dependent in the reference (deps/spfa/src/main.cpp:65-90) and chosen by a
    fixed direction order in the kernels; among EQUAL-COST parent chains the extracted waypoints can differ (SURVEY.md §7 "SSSP
    ties").  With the flag, a step whose spfa path has the same end points but other intermediate waypoints returns
    {"tie": True} instead of failing; the caller stops comparing that env (trajectories only agree statistically from there)."""
    obs, rew, term, trunc, info = step_out
    oobs, orr, oterm, otrunc, oinfo = ora.step(action)
    to_np = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    gb = env.get_bodies(e)
    ob = ora.space.state()
    dbg = env.get_step_debug(e)
    if allow_path_ties and dbg["need_spfa"] and len(ora.last_path) >= 3:
        mine = np.asarray(dbg["path"])[: dbg["n_path"], :2]
        theirs = np.asarray([p[:2] for p in ora.last_path], dtype=np.float64)
        if mine.shape != theirs.shape or np.abs(mine - theirs).max() > 1e-9:
            assert np.abs(mine[0] - theirs[0]).max() <= 1e-9 and dbg["status"] == 0      # same start; both are shortest grid paths
            return {"tie": True}
    nb = len(ora.box_bodies)
    sd = float(np.abs(gb[:1 + nb, :3] - ob[:1 + nb, :3]).max())
    assert sd <= state_tol, f"pose differs from the oracle: {sd}"
    assert abs(float(to_np(rew)[e]) - orr) <= 1e-6 * max(1.0, abs(orr)), (float(to_np(rew)[e]), orr)
    assert bool(to_np(term)[e]) == bool(oterm) and bool(to_np(trunc)[e]) == bool(otrunc)
    assert dbg["n_path"] == len(ora.last_path) and dbg["status"] == 0
    assert dbg["substeps"] == 1000 + ora.n_substeps, "different number of physics substeps"
    i = {k: to_np(v) for k, v in info.items() if k not in ("goal_positions", "boundary")}
    assert int(i["box_count"][e]) == oinfo["box_count"] and int(i["t"][e]) == ora.t
    assert abs(float(i["total_work"][e]) - oinfo["total_work"]) <= 1e-7
    assert abs(float(i["diff_reward"][e]) - oinfo["diff_reward"]) <= 1e-7
    assert abs(float(i["box_completed_reward"][e]) - oinfo["box_completed_reward"]) <= 1e-12
    assert abs(float(i["collision reward"][e]) - oinfo["collision reward"]) <= 1e-7
    assert list(i["box_completed_statuses"][e][:nb].astype(bool)) == list(oinfo["box_completed_statuses"])
    assert np.allclose(i["obs"][e][:nb], np.stack(oinfo["obs"]), atol=1e-6)
    assert np.allclose(i["state"][e], np.array(oinfo["state"]), atol=0.011)
    nd = int((to_np(obs)[e] != oobs).sum())
    assert nd <= max_obs_diff, f"{nd} observation bytes differ from the oracle"
    return {"state": sd, "obs": nd}
