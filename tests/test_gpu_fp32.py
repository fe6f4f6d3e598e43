"""The fp32-state build of the same kernel sources (libbenchpush_b200_f32.so: real = float for body state, contacts and
solver; motion bounds, gates, accumulated work and outputs stay double) against the fp64 oracle.

Stated tolerances (SURVEY.md §8d asks 1e-3 m / 1e-2 rad per env step on active floes, rasters bit-exact on identical states,
episode scores statistically indistinguishable over 200 seeded episodes):
  * one env step from the reset state: pose within 1e-3 m / 1e-2 rad of the fp64 oracle (measured <= 3.8e-4 m / 1.4e-3 rad at
    concentration 0.5, 7.9e-6 m at 0.1 — below the solver-ORDER ambiguity of the reference itself, <= 5.4e-3 m, see
    tests/test_oracle_cpu.py); the first five steps within 2e-3 m / 1e-2 rad;
  * the observation is the bit-exact raster of the state the env is in (oracle rasteriser fed with the env's own poses);
  * the 200 seeded episodes of BASELINE config 1 on the reference's own ice fields: success rate inside the binomial 95 %
    interval, KS p > 0.05 on efficiency and effort against the fp64 canonical-order golden.
Throughput of this build: profiles/r02c_fp32_vs_fp64.txt (+25 % on the batched kernel)."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import FIXTURES, GOLD, ROOT, SHIP_VERTS, fixture_conc
from benchpush_b200.layouts import unpack_trials
from oracle.ship_ice_oracle import ShipIceOracleEnv, total_work_done

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _env(p, conc, n=1, **kw):
    from benchpush_b200.envs import ShipIceVecEnv
    return ShipIceVecEnv(num_envs=n, cfg={"concentration": conc}, layouts=p, device=0, precision="fp32", **kw)


@pytest.mark.parametrize("name", FIXTURES)
def test_fp32_transitions_within_tolerance_and_rasters_exact(name, golden_layouts):
    from oracle import raster_oracle as ro
    p = golden_layouts[name]
    conc = fixture_conc(name)
    env = _env(p, conc)
    obs, info = env.reset()
    ora = ShipIceOracleEnv({"concentration": conc}, experiment=unpack_trials(p), order_mode=1)
    ora.reset()
    d0 = np.abs(env.get_bodies(0)[:, :3] - ora.space.state()[:, :3])
    assert d0[:, :2].max() <= 2e-3 and d0[:, 2].max() <= 1e-2                     # 1000 settle substeps in fp32
    occ = ro.OccupancyGrid(grid_width=1 / 25, grid_height=1 / 25, map_width=12, map_height=40, local_width=6, local_height=6, meter_to_pixel_scale=25)
    rng = np.random.default_rng(1)
    prev = env.get_obstacles(0)
    for k in range(5):
        a = np.float32(rng.uniform(-1, 1))
        obs, rew, term, trunc, info = env.step(np.array([a], dtype=np.float32))
        ora.step(float(a), want_obs=False)
        g, o = env.get_bodies(0), ora.space.state()
        lim_p, lim_a = (1e-3, 1e-2) if k == 0 else (2e-3, 1e-2)
        assert np.abs(g[:, :2] - o[:, :2]).max() <= lim_p and np.abs(g[:, 2] - o[:, 2]).max() <= lim_a, (k, np.abs(g[:, :3] - o[:, :3]).max(axis=0))
        # raster = the oracle's rasteriser on the state the env is in (poses as exported, float -> double is exact)
        pose = tuple(float(x) for x in info["pose"][0].cpu().numpy())
        cur = env.get_obstacles(0)
        raw = occ.compute_occ_img(cur, ice_binary_w=300, ice_binary_h=1000, local_range=[12, 12], ship_state=pose)
        ch = [occ.ego_view_footprint(pose, SHIP_VERTS, 2), occ.ego_view_goal_dist_transform(9.0, pose, 2),
              occ.ego_view_orientation_map(pose, [1., 0.], [-1., 0.], 2), occ.ego_view_obstacle_map(raw, pose, 2)]
        want = (np.stack([np.array(c) for c in ch]) * 255).astype(np.uint8)
        assert np.array_equal(obs[0].cpu().numpy(), want), int((obs[0].cpu().numpy() != want).sum())
        # work of the step recomputed in fp64 from the env's own before / after vertices
        w = total_work_done(prev, cur)
        np.testing.assert_allclose(-float(info["collision reward"][0]), w, rtol=1e-7, atol=1e-12)
        prev = cur
    env.close()


def test_fp32_resident_step_equals_fp32_batched_step(monkeypatch):
    from benchpush_b200.layouts import generate_experiment, pack_trials
    packed = pack_trials(generate_experiment(4, 0.5, seed=5))
    envs = {}
    for impl in ("batch", "res"):
        monkeypatch.setenv("BP_STEP_IMPL", impl)
        envs[impl] = _env(packed, 0.5, n=40, auto_reset=True)
        envs[impl].reset()
    monkeypatch.delenv("BP_STEP_IMPL")
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for k in range(12):
        a = torch.rand(40, generator=g, device="cuda") * 2 - 1
        ra, rb = envs["batch"].step(a), envs["res"].step(a)
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]), k
        for e in (0, 9, 39):
            assert np.array_equal(envs["batch"].get_bodies(e), envs["res"].get_bodies(e)), (k, e)
    for v in envs.values():
        v.close()


def test_fp32_200_seeded_episodes_on_the_reference_ice_fields():
    import make_golden_200ep as g
    from benchpush_b200.layouts import load_packed
    from benchpush_b200.metrics import ShipIceMetric
    from tests.test_episodes_200 import _run_batch, _stat_agree
    gold = np.load(os.path.join(GOLD, "shipice_200ep_refdata.npz"))
    E = g.N_EP
    env = _env(load_packed(g.REF_PACK), g.CONC, n=E, first_trial=np.arange(E) % 100)
    met = ShipIceMetric("gpu-fp32", ship_mass=env.cfg.ship.mass, goal=env.goal)
    steps, succ, eff, effort, rew = _run_batch(env, met, gold["actions"], g.MAX_STEPS)
    env.close()
    assert np.mean(steps == gold["canonical_steps"]) >= 0.95                      # measured: every episode has the oracle's length
    for tag in ("canonical", "libm"):
        _stat_agree(succ, gold["%s_success" % tag], eff, gold["%s_efficiency" % tag], effort, gold["%s_effort" % tag])
    assert abs(float(np.mean(effort)) - float(np.mean(gold["canonical_effort"]))) < 5e-3
