"""Shared scenario helpers of the box-delivery tests (CPU emulation and GPU)."""
import numpy as np

DELIVERY_LAYOUT = {"start": (1.5, 1.75, 0.0), "boxes": [[3.0, 1.75, 0.3], [2.9, 0.9, 1.0], [-2.0, -1.0, 0.5], [0.0, 0.0, 0.2]]}
DELIVERY_CFG = {"agent": {"action_type": "heading"}, "boxes": {"num_boxes_small": 4}}


def compare_step(env, ora, e, action_env, action_ora, step_out=None, state_tol=1e-7, max_obs_diff=4):
    """Steps the oracle with `action_ora`; the vector env has already been stepped (step_out) or is stepped here for
    env `e` only when num_envs == 1.  Returns a dict of the largest differences."""
    if step_out is None:
        step_out = env.step(np.array([action_env], dtype=np.float32))
    obs, rew, term, trunc, info = step_out
    oobs, orr, oterm, otrunc, oinfo = ora.step(action_ora)
    to_np = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    gb = env.get_bodies(e)
    ob = ora.space.state()
    dbg = env.get_step_debug(e)
    alive = gb[1:, 6] == 1
    assert int(alive.sum()) == len(ora.box_bodies), "different boxes left"
    sd = float(np.abs(gb[0, :3] - ob[0, :3]).max())
    bd = float(np.abs(gb[1:][alive][:, :3] - ob[list(ora.box_bodies)][:, :3]).max()) if len(ora.box_bodies) else 0.0
    assert max(sd, bd) <= state_tol, f"pose differs from the oracle: robot {sd}, boxes {bd}"
    assert abs(float(to_np(rew)[e]) - orr) <= 1e-6 * max(1.0, abs(orr)), (float(to_np(rew)[e]), orr)
    assert bool(to_np(term)[e]) == bool(oterm) and bool(to_np(trunc)[e]) == bool(otrunc)
    assert dbg["n_path"] == len(ora.last_path) and dbg["status"] == 0
    assert np.allclose(dbg["path"][:, :2], np.array([p[:2] for p in ora.last_path], dtype=np.float64), atol=1e-7)
    assert dbg["robot_hit"] == int(ora.space.robot_hit() == 1)
    assert dbg["substeps"] == 1000 + ora.n_substeps, "different number of physics substeps"
    i = {k: to_np(v) for k, v in info.items() if k not in ("goal_positions",)}
    assert int(i["cumulative_boxes"][e]) == oinfo["cumulative_boxes"] and int(i["inactivity"][e]) == oinfo["inactivity"]
    assert abs(float(i["total_work"][e]) - oinfo["total_work"]) <= 1e-7
    assert abs(float(i["cumulative_distance"][e]) - oinfo["cumulative_distance"]) <= 1e-7
    assert np.allclose(i["state"][e], np.array(oinfo["state"]), atol=0.011)          # 2-dp rounding may flip on a 1e-9 difference
    assert list(i["box_completed_statuses"][e].astype(bool)) == list(oinfo["box_completed_statuses"])
    nleft = int(i["n_boxes_left"][e])
    assert nleft == len(oinfo["obs"])
    if nleft:
        assert np.allclose(i["obs"][e][:nleft], np.stack(oinfo["obs"]), atol=1e-6)
    nd = 0
    if oobs is not None:
        nd = int((to_np(obs)[e] != oobs).sum())
        assert nd <= max_obs_diff, f"{nd} observation bytes differ from the oracle"
    return {"state": max(sd, bd), "obs": nd, "spfa": dbg["need_spfa"]}
