"""The C-ABI library loads on a CPU-only box and exports every symbol include/benchpush_b200.h
declares; host-side packing / config logic.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT
from benchpush_b200 import _lib
from benchpush_b200.layouts import generate_experiment, pack_trials, unpack_trials
from benchpush_b200.utils import DotDict, merge_cfg


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "benchpush_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTS)


def test_library_reports_no_device_without_cuda():
    L = _lib.lib()
    assert L.bp_version() >= 100
    import torch
    if not torch.cuda.is_available():
        assert L.bp_device_count() == 0
        cfg = _lib.ShipIceConfig()
        L.bp_shipice_default_config(ctypes.byref(cfg))
        h = ctypes.c_void_p()
        rc = L.bp_shipice_create(ctypes.byref(cfg), ctypes.byref(h))
        assert rc == -2                                   # BP_ERR_CUDA: fails loudly, no CPU fallback
        with pytest.raises(_lib.BenchPushError):
            _lib.check(rc)


def test_default_config_matches_reference_defaults():
    cfg = _lib.ShipIceConfig()
    _lib.lib().bp_shipice_default_config(ctypes.byref(cfg))
    assert (cfg.dt, cfg.steps, cfg.iterations, cfg.damping) == (0.8, 400, 10, 0.0)
    assert (cfg.target_speed, cfg.goal_y, cfg.map_width, cfg.map_height, cfg.m_to_pix_scale) == (0.3, 9.0, 12.0, 40.0, 25)
    assert cfg.n_ship_verts == 17 and cfg.settle_steps == 1000
    y = DotDict.load_from_file(os.path.join(ROOT, "benchpush_b200", "configs", "ship_ice.yaml"))
    assert np.allclose(np.array(y.ship.vertices).ravel(), np.array(cfg.ship_verts[:34]))
    assert y.sim.steps == 400 and y.sim.damping == 0 and y.occ.m_to_pix_scale == 25 and y.goal_y == 9


def test_cfg_merge_is_one_level_deep():
    base = DotDict.to_dot_dict({"a": 1, "sim": {"steps": 400, "damping": 0}, "ship": {"mass": 1}})
    out = merge_cfg(base, {"a": 2, "sim": {"steps": 10}, "new": {"k": 3}})
    assert out.a == 2 and out.sim.steps == 10 and out.sim.damping == 0 and out.new.k == 3 and out.ship.mass == 1
    assert base.sim.steps == 400


def test_pack_unpack_roundtrip_and_generator_statistics():
    trials = generate_experiment(2, 0.5, seed=11)
    p = pack_trials(trials)
    assert p.n_layouts == 2 and p.verts.shape[2:] == (20, 2)
    back = unpack_trials(p)
    for t, b in zip(trials, back):
        assert len(t["obstacles"]) == len(b["obstacles"])
        assert np.array_equal(t["obstacles"][3]["vertices"], b["obstacles"][3]["vertices"])
    n = p.n_floes
    assert np.all(n > 500) and np.all(n <= 653)               # reference at 50 %: 589..653 floes
    nv = p.n_verts[0, : n[0]]
    assert nv.min() >= 10 and nv.max() <= 20
    # determinism
    p2 = pack_trials(generate_experiment(2, 0.5, seed=11))
    assert np.array_equal(p.verts, p2.verts)


def test_create_rejects_bad_arguments_before_touching_cuda():
    """Error behaviour of the C ABI (include/benchpush_b200.h: negative status + bp_last_error(), nothing thrown, nothing
    allocated): argument checks come first, so they are observable on a machine without a GPU.  The reference's counterpart is a
    Python exception from the env constructor; the ctypes mirror turns the status into BenchPushError."""
    L = _lib.lib()
    L.bp_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    INVALID = -1
    # ship-ice
    cfg = _lib.ShipIceConfig(); L.bp_shipice_default_config(ctypes.byref(cfg))
    assert L.bp_shipice_create(None, ctypes.byref(h)) == INVALID and b"null" in L.bp_last_error()
    assert L.bp_shipice_create(ctypes.byref(cfg), None) == INVALID
    for field, bad in (("n_envs", 0), ("n_envs", -3), ("max_floes", 0), ("max_floes", 60001), ("n_ship_verts", 2), ("n_ship_verts", 33),
                       ("steps", 0), ("iterations", 0), ("dt", 0.0)):
        c = _lib.ShipIceConfig(); L.bp_shipice_default_config(ctypes.byref(c))
        if c.n_envs <= 0: c.n_envs = 4
        if c.max_floes <= 0: c.max_floes = 100
        setattr(c, field, bad)
        assert L.bp_shipice_create(ctypes.byref(c), ctypes.byref(h)) == INVALID, field
        assert L.bp_last_error(), field
        with pytest.raises(_lib.BenchPushError):
            _lib.check(INVALID)
    # maze-NAMO
    m = _lib.MazeConfig(); L.bp_maze_default_config(ctypes.byref(m), 1)
    assert L.bp_maze_create(None, ctypes.byref(h)) == INVALID
    for field, bad in (("n_envs", 0), ("max_boxes", -1), ("max_boxes", 4097), ("n_walls", -1), ("n_robot_verts", 2), ("steps", 0), ("m_to_pix_scale", 0)):
        c = _lib.MazeConfig(); L.bp_maze_default_config(ctypes.byref(c), 1)
        if c.n_envs <= 0: c.n_envs = 4
        setattr(c, field, bad)
        assert L.bp_maze_create(ctypes.byref(c), ctypes.byref(h)) == INVALID, field
    # room envs
    from benchpush_b200.envs.box_delivery import BoxDelConfig, bind
    bind(L)
    for field, bad in (("n_envs", 0), ("num_boxes", 0), ("num_boxes", 10 ** 6), ("action_type", 3), ("action_type", -1)):
        c = BoxDelConfig(); L.bp_boxdel_default_config(ctypes.byref(c))
        if c.n_envs <= 0: c.n_envs = 4
        if c.num_boxes <= 0: c.num_boxes = 10
        setattr(c, field, bad)
        assert L.bp_boxdel_create(ctypes.byref(c), ctypes.byref(h)) == INVALID, field
    assert L.bp_boxdel_create(None, ctypes.byref(h)) == INVALID
    assert not h.value                                             # no handle was ever produced
