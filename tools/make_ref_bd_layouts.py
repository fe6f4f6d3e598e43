"""Golden box-delivery episode layouts from the REFERENCE ITSELF (tests/golden/ref_bd_layouts.npz).

Imports benchpush/environments/box_delivery/box_delivery_env.py from /root/reference with stub modules for the absent
packages (gymnasium, pymunk, pygame, matplotlib, skimage, shapely: tools/make_ref_fixtures.install_stubs plus the ones
below; none of them is called on this path) and runs the reference's own
    BoxDeliveryEnv.get_random_robot_start   (box_delivery_env.py:313-320)
    BoxDeliveryEnv.generate_boundary        (:326-507: random columns / horizontal divider, corners)
    BoxDeliveryEnv.generate_boxes           (:509-559)
in the order init_box_delivery_env calls them (:241-252) on an instance made with ``__new__`` that carries exactly the
attributes those three methods read.  tests/test_ref_pins_cpu.py holds benchpush_b200.layouts.sample_box_delivery_layout
and the oracle's sampler to the result.  Run in the build container only:  python tools/make_ref_bd_layouts.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "ref_bd_layouts.npz")
CONFIGS = ["small_empty", "small_columns", "large_columns", "large_divider"]
SEEDS = list(range(12))


def main():
    from benchpush_b200.envs.box_delivery import DEFAULT_CFG      # (pulls torch: import before the stubs go in)
    from benchpush_b200.utils import merge_cfg
    import make_ref_fixtures as mrf
    mrf.install_stubs()

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        def ga(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()
        m.__getattr__ = ga
        sys.modules[name] = m
        return m

    class Env:
        pass

    stub("gymnasium.envs.registration", register=lambda *a, **k: None)
    stub("gymnasium.envs", registration=sys.modules["gymnasium.envs.registration"])
    stub("gymnasium", Env=Env, spaces=stub("gymnasium.spaces"), envs=sys.modules["gymnasium.envs"])
    stub("pymunk", Vec2d=tuple, pygame_util=stub("pymunk.pygame_util"))
    stub("pygame")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))      # the reference's compiled spfa module
    sys.path.insert(0, REF)
    from benchpush.environments.box_delivery.box_delivery_env import BoxDeliveryEnv
    from benchpush.common.utils.utils import DotDict

    out = {"configs": np.array(CONFIGS), "seeds": np.array(SEEDS)}
    for oc in CONFIGS:
        cfgd = merge_cfg(DEFAULT_CFG, {"env": {"obstacle_config": oc}})
        cfg = DotDict.to_dot_dict(cfgd) if hasattr(DotDict, "to_dot_dict") else DotDict(cfgd)
        for seed in SEEDS:
            e = BoxDeliveryEnv.__new__(BoxDeliveryEnv)
            e.cfg = cfg
            e.random_state = np.random.RandomState(seed)
            small = oc.split("_")[0] == "small"
            e.room_length = cfg.env.room_length
            e.room_width = cfg.env.room_width_small if small else cfg.env.room_width_large
            e.num_boxes = cfg.boxes.num_boxes_small if small else cfg.boxes.num_boxes_large
            e.wall_thickness = cfg.env.wall_thickness
            e.robot_info = cfg.agent
            e.robot_radius = ((e.robot_info.length ** 2 + e.robot_info.width ** 2) ** 0.5 / 2) * 1.2      # box_delivery_env.py:130
            e.receptacle_position, e.receptacle_size = e.get_receptacle_position_and_size()
            e.start = e.get_random_robot_start() if cfg.agent.random_start else (5, 1.5, np.pi * 3 / 2)
            e.robot_info["start_pos"] = e.start
            e.boundary_dicts = e.generate_boundary()
            boxes = e.generate_boxes()
            key = "%s_%d_" % (oc, seed)
            out[key + "start"] = np.array(e.robot_info["start_pos"], dtype=np.float64)
            out[key + "columns"] = np.array([b["position"] for b in e.boundary_dicts if b["type"] == "column"], dtype=np.float64).reshape(-1, 2)
            out[key + "divider"] = np.array([b["position"] for b in e.boundary_dicts if b["type"] == "divider"], dtype=np.float64).reshape(-1, 2)
            out[key + "boxes"] = np.array([[b["position"][0], b["position"][1], b["heading"]] for b in boxes], dtype=np.float64)
            out[key + "types"] = np.array([b["type"] for b in e.boundary_dicts])
            out[key + "corners"] = np.array([[b["position"][0], b["position"][1], b["heading"]] for b in e.boundary_dicts if b["type"] == "corner"], dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
