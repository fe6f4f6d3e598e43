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


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    make_layouts()
    make_spfa()
    make_pins()
    print("golden fixtures written to", GOLD)
