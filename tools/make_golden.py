"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists).  Nothing under tests/ or bench.py reads /root/reference at run time.

  ice_cXX_tY.npz        packed ice-field layouts taken from the reference pickles
                        (ship_ice_nav/ice_environments/experiments_XX_100_r06_d40x12.pk) — INPUT fixtures
  spfa_golden.npz       maps, sources and the outputs of the reference's own spfa
                        (deps/spfa/src/main.cpp compiled in place -> oracle/_ref) — PINNED outputs
  shipice_oracle_pins.npz   trajectory pins produced by the (unpinned) physics oracle, to detect drift
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from benchpush_b200.layouts import pack_trials, save_packed, unpack_trials  # noqa: E402
from tools.ref_data import load_experiment  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FIXTURES = [(0.1, 0), (0.1, 7), (0.3, 7), (0.5, 0)]


def make_layouts():
    for conc, trial in FIXTURES:
        exp = load_experiment(conc)
        p = pack_trials([exp[trial]])
        save_packed(os.path.join(GOLD, "ice_c%02d_t%d.npz" % (int(conc * 100), trial)), p)


def make_spfa():
    import spfa as ref_spfa
    rng = np.random.default_rng(1234)
    maps, srcs, dists, parents, shapes = [], [], [], [], []
    cases = [(1, 1, 0.0), (1, 7, 0.0), (5, 1, 0.0), (7, 9, 0.25), (16, 16, 0.4), (33, 47, 0.3), (64, 64, 0.3), (96, 128, 0.35), (107, 135, 0.1)]
    for (H, W, p) in cases:
        for rep in range(2):
            m = rng.random((H, W)) > p
            if H > 8:
                for _ in range(4):
                    i, j = rng.integers(0, H), rng.integers(0, W)
                    m[i:i + max(1, H // 5), j:j + max(1, W // 12)] = False
            src = (int(rng.integers(0, H)), int(rng.integers(0, W)))
            if rep == 0:
                m[src] = True
            d, par = ref_spfa.spfa(m, src)
            maps.append(m.astype(np.uint8).ravel()); srcs.append(src); dists.append(d.ravel()); parents.append(par.ravel()); shapes.append((H, W))
    np.savez_compressed(os.path.join(GOLD, "spfa_golden.npz"), shapes=np.array(shapes, np.int32), sources=np.array(srcs, np.int32),
                        maps=np.concatenate(maps), dists=np.concatenate(dists), parents=np.concatenate(parents))


def make_pins():
    from benchpush_b200.layouts import load_packed
    from oracle.ship_ice_oracle import ShipIceOracleEnv
    out = {}
    for conc, trial in FIXTURES:
        name = "c%02d_t%d" % (int(conc * 100), trial)
        p = load_packed(os.path.join(GOLD, "ice_%s.npz" % name))
        for order in (0, 1):
            env = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=order)
            obs, info = env.reset()
            rng = np.random.default_rng(42)
            nsteps = 6 if conc >= 0.5 else 10
            actions = rng.uniform(-1, 1, nsteps).astype(np.float32).astype(np.float64)   # exactly representable in the f32 action space
            rows, obs_sums = [], [obs.astype(np.int64).reshape(4, -1).sum(1)]
            for a in actions:
                obs, r, term, _, info = env.step(float(a))
                st = env.space.state()
                rows.append([st[-1, 0], st[-1, 1], st[-1, 2], r, info["total_work"], st[:-1, 0].sum(), st[:-1, 1].sum(), st[:-1, 2].sum(), float(term)])
                obs_sums.append(obs.astype(np.int64).reshape(4, -1).sum(1))
            out["%s_o%d_actions" % (name, order)] = actions
            out["%s_o%d_rows" % (name, order)] = np.array(rows)
            out["%s_o%d_obs_sums" % (name, order)] = np.array(obs_sums)
            if order == 1:
                out["%s_final_state" % name] = env.space.state()
                out["%s_final_obs" % name] = obs
    np.savez_compressed(os.path.join(GOLD, "shipice_oracle_pins.npz"), **out)


MAZE_LAYOUTS = [[[11.3, 6.0], [11.0, 9.5], [9.4, 12.4], [5.0, 12.6], [3.6, 8.0]],      # obstacles on the way to the goal
                [[13.6, 5.0], [13.9, 8.0], [8.8, 11.2], [6.5, 11.5], [3.8, 6.0]]]      # obstacles next to the walls
MAZE_WAYPOINTS = [(11.25, 12.2), (3.75, 12.4), (3.75, 3.75)]


def maze_scripted_action(pose, wi, k, ep):
    """Heading controller towards the waypoints (test driver only): returns (action, waypoint index)."""
    import math
    px, py, pa = pose
    if ep == 1 and k < 40:
        tx, ty = 14.0, 8.0                      # push an obstacle into the right wall, then hit the wall
    else:
        tx, ty = MAZE_WAYPOINTS[wi]
        if math.hypot(tx - px, ty - py) < 1.0 and wi < len(MAZE_WAYPOINTS) - 1:
            wi += 1
            tx, ty = MAZE_WAYPOINTS[wi]
    err = math.atan2(ty - py, tx - px) - pa
    err = (err + math.pi) % (2 * math.pi) - math.pi
    a = float(np.float32(np.clip(err / (0.8 * (math.pi / 2) / 15), -1, 1)))
    return a, wi


def make_maze():
    """maze_golden.npz: (a) PINNED third-party outputs — scipy.ndimage.rotate(order=1) and scipy.special.sindg/cosdg
    of the scipy in this image; (b) trajectory pins of the (unpinned) maze oracle with the portable cos/sin
    (bit-reproducible on every machine): recorded actions, per-step rows, final states and observations."""
    from scipy import special
    from scipy.ndimage import rotate
    from benchpush_b200.layouts import generate_maze_layouts
    from oracle import cpo
    from oracle.maze_oracle import MazeNamoOracleEnv
    out = {}
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-720, 720, 4000), np.arange(-360, 361, 15.0)])
    out["dg_x"] = xs; out["dg_sin"] = special.sindg(xs); out["dg_cos"] = special.cosdg(xs)
    for k, (n, binary, cval) in enumerate([(64, True, 0.0), (48, False, 1.0), (96, True, 0.0)]):
        img = (rng.random((n, n)) < 0.35).astype(np.float64) if binary else rng.random((n, n))
        ang = float(rng.uniform(-200, 200))
        out["rot%d_in" % k] = img.astype(np.float32) if False else img
        out["rot%d_angle" % k] = np.array([ang, cval])
        out["rot%d_out" % k] = rotate(img, ang, reshape=False, cval=cval, order=1)
    layouts = [list(l) for l in MAZE_LAYOUTS] + generate_maze_layouts(2, seed=5)
    out["layouts"] = np.array(layouts, dtype=np.float64)
    cpo.set_portable_trig(True)
    try:
        env = MazeNamoOracleEnv(layouts=layouts, order_mode=1)
        for ep in range(len(layouts)):
            obs, info = env.reset()
            out["ep%d_reset_obs" % ep] = obs
            out["ep%d_reset_state" % ep] = env.space.state()
            wi = 0
            arng = np.random.default_rng(100 + ep)
            actions, rows = [], []
            for k in range(260):
                if ep < 2:
                    a, wi = maze_scripted_action(env.robot_pose(), wi, k, ep)
                else:
                    a = float(np.float32(arng.uniform(-1, 1)))
                obs, r, term, _, info = env.step(a)
                st = env.space.state()
                actions.append(a)
                rows.append([st[11, 0], st[11, 1], st[11, 2], r, info["total_work"], float(term), float(info["trial_success"]),
                             info["dist increment reward"], st[6:11, 0].sum(), st[6:11, 1].sum(), st[6:11, 2].sum()])
                if k in (10, 60):
                    out["ep%d_obs_k%d" % (ep, k)] = obs
                if term:
                    break
            out["ep%d_actions" % ep] = np.array(actions)
            out["ep%d_rows" % ep] = np.array(rows)
            out["ep%d_final_state" % ep] = env.space.state()
            out["ep%d_final_obs" % ep] = obs
    finally:
        cpo.set_portable_trig(False)
    np.savez_compressed(os.path.join(GOLD, "maze_golden.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "maze":
        make_maze()
        print("maze fixtures written to", GOLD)
        sys.exit(0)
    make_layouts()
    make_spfa()
    make_pins()
    make_maze()
    print("golden fixtures written to", GOLD)
