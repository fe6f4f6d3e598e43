"""Multi-rank host logic on CPU: env sharding by global env id and the episode-metric reduction
(the only collective of the path), world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from benchpush_b200.parallel import shard_range
from benchpush_b200.metrics import ShipIceMetric


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    m = ShipIceMetric("t", ship_mass=1, goal=(0, 9))
    n = hi - lo
    info0 = {"state": torch.tensor([[5.0, 1.0, 1.57]] * n, dtype=torch.float64)}
    m.reset(info0)
    info1 = {"state": torch.tensor([[5.0, 9.0, 1.57]] * n, dtype=torch.float64), "total_work": torch.full((n,), 2.0, dtype=torch.float64),
             "trial_success": torch.tensor([(lo + i) % 2 == 0 for i in range(n)])}
    m.update(info1, torch.full((n,), 3.0, dtype=torch.float64), torch.ones(n, dtype=torch.bool))
    q.put((rank, lo, hi, m.reduce()))
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n, w in [(10, 2), (65536, 8), (7, 4), (3, 8)]:
        cover = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            cover += list(range(lo, hi))
        assert cover == list(range(n))


def test_metric_reduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    res.sort()
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 5, 5, 10)
    for _, _, _, red in res:
        assert red["episodes"] == 10
        np.testing.assert_allclose(red["success_rate"], 0.5)
        np.testing.assert_allclose(red["efficiency_mean"], 0.5 * (8.0 / 8.0))
        np.testing.assert_allclose(red["effort_mean"], 8.0 / (8.0 + 2.0))
        np.testing.assert_allclose(red["reward_mean"], 3.0)


def _worker_rooms(rank, world, port, q):
    """box-delivery / area-clearing shards (CPU emulation backend): layouts keyed by the global env id, TaskDrivenMetric reduced over ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from benchpush_b200.envs.area_clearing import AreaClearingVecEnv
    from benchpush_b200.metrics import TaskDrivenMetric
    from tests.emu import emu_bd
    lo, hi = shard_range(4, rank, world)
    env = emu_bd.EmuBoxDeliveryVecEnv(num_envs=hi - lo, cfg={"agent": {"action_type": "heading"}}, env_offset=lo)
    ac = emu_bd.EmuAreaClearingVecEnv(num_envs=hi - lo, env_offset=lo)
    m = TaskDrivenMetric("t", robot_mass=1)
    obs, info = env.reset()
    m.reset(info)
    acts = np.linspace(-0.8, 0.8, 4).astype(np.float32)[lo:hi]
    obs, rew, term, trunc, info = env.step(acts)
    m.update(info, rew, eps_complete=True)
    n = hi - lo            # the pool holds several episodes per env: entry f * n + e = (env lo + e, episode f)
    q.put((rank, [[l["start"] for l in env.layouts[f * n:(f + 1) * n]] for f in range(len(env.layouts) // n)],
           [l["start"] for l in ac.layouts[:n]], env.get_bodies(0)[0, :3].tolist(), m.reduce()))
    env.close(); ac.close()
    dist.destroy_process_group()


def test_room_envs_shard_by_global_env_id_world2_gloo():
    from benchpush_b200.envs.box_delivery import BoxDeliveryVecEnv
    from tests.emu import emu_bd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_rooms, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=300) for _ in ps)
    [p.join(timeout=60) for p in ps]
    whole = emu_bd.EmuBoxDeliveryVecEnv(num_envs=4, cfg={"agent": {"action_type": "heading"}})
    F = len(whole.layouts) // 4
    assert F == len(res[0][1]) == len(res[1][1]) and F > 1
    for f in range(F):       # same layouts as one rank with all envs, for every episode of the pool
        assert res[0][1][f] + res[1][1][f] == [l["start"] for l in whole.layouts[4 * f:4 * f + 4]]
    assert len(set(l["start"] for l in whole.layouts)) == 4 * F                   # and every (env, episode) has its own layout
    assert len(set(res[0][2] + res[1][2])) == 4
    whole.reset()
    whole.step(np.linspace(-0.8, 0.8, 4).astype(np.float32))
    assert whole.get_bodies(2)[0, :3].tolist() == res[1][3]                        # env 2 = first env of rank 1: same trajectory
    assert res[0][4]["episodes"] == 4 and res[0][4] == res[1][4]
    whole.close()
