"""Efficiency / effort scoring (reference benchpush/common/metrics/base_metric.py:7-105 and
ship_ice_metric.py:5-72), batched over environments.

The reference keeps one metric object per (single) env and appends one score per finished episode to
``rewards``, ``efficiency_scores``, ``effort_scores``.  ``ShipIceMetric`` here keeps the same
attribute names and formulas; ``update`` consumes the batched ``info`` of ``ShipIceVecEnv.step`` and
appends one entry for every env whose episode ended in that step.  Plotting is out of scope.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class BaseMetric(ABC):
    def __init__(self, alg_name) -> None:
        self.rewards = []
        self.efficiency_scores = []
        self.effort_scores = []
        self.success_rates = []
        self.alg_name = alg_name

    @abstractmethod
    def update(self, info, reward, eps_complete=False):
        ...

    @abstractmethod
    def reset(self, info, mask=None):
        ...


class ShipIceMetric(BaseMetric):
    """efficiency = 1_success * L / l_0 with L = goal_y - y_start (ship_ice_metric.py:26-34, 72);
    effort = m_0 l_0 / (m_0 l_0 + sum_i m_i l_i) with sum_i m_i l_i = info['total_work'] (:37-43).
    The ship path length l_0 integrates the 2-decimal-rounded ``info['state']`` (:52-55)."""

    def __init__(self, alg_name, ship_mass, goal, num_envs=None, device=None) -> None:
        super().__init__(alg_name=alg_name)
        self.ship_mass = float(ship_mass)
        self.goal_line = float(goal[1])
        self._n = num_envs
        self._dev = device
        self.successes = []
        self._alloc_done = False

    def _alloc(self, info):
        st = info["state"]
        self._n, self._dev = st.shape[0], st.device
        z = lambda: torch.zeros(self._n, dtype=torch.float64, device=self._dev)
        self.eps_reward, self.total_mass_dist, self.total_ship_dist, self.L = z(), z(), z(), z()
        self.trial_success = torch.zeros(self._n, dtype=torch.bool, device=self._dev)
        self.ship_state = torch.zeros((self._n, 3), dtype=torch.float64, device=self._dev)
        self._alloc_done = True

    def reset(self, info, mask=None):
        if not self._alloc_done:
            self._alloc(info)
        m = torch.ones(self._n, dtype=torch.bool, device=self._dev) if mask is None else mask.bool()
        zero = torch.zeros((), dtype=torch.float64, device=self._dev)
        self.eps_reward = torch.where(m, zero, self.eps_reward)
        self.total_mass_dist = torch.where(m, zero, self.total_mass_dist)
        self.total_ship_dist = torch.where(m, zero, self.total_ship_dist)
        self.trial_success = self.trial_success & ~m
        self.ship_state = torch.where(m[:, None], info["state"], self.ship_state)
        self.L = torch.where(m, self.goal_line - info["state"][:, 1], self.L)

    def compute_efficiency_score(self):
        eff = self.L / self.total_ship_dist
        return torch.where(self.trial_success, eff, torch.zeros_like(eff))

    def compute_effort_score(self):
        ml = self.ship_mass * self.total_ship_dist
        return ml / (ml + self.total_mass_dist)

    def update(self, info, reward, eps_complete=False):
        """eps_complete: bool tensor [E] (``terminated | truncated``) or a python bool applied to all."""
        self.eps_reward = self.eps_reward + reward
        self.total_mass_dist = info["total_work"].clone()
        self.trial_success = info["trial_success"].clone()
        st = info["state"]
        self.total_ship_dist = self.total_ship_dist + torch.linalg.norm(self.ship_state[:, :2] - st[:, :2], dim=1)
        self.ship_state = st.clone()
        done = eps_complete if torch.is_tensor(eps_complete) else torch.full((self._n,), bool(eps_complete), device=self._dev)
        if bool(done.any()):
            idx = torch.nonzero(done).flatten()
            self.rewards.extend(self.eps_reward[idx].tolist())
            self.efficiency_scores.extend(self.compute_efficiency_score()[idx].tolist())
            self.effort_scores.extend(self.compute_effort_score()[idx].tolist())
            self.successes.extend(self.trial_success[idx].tolist())
        return done

    def reduce(self):
        """Episode-metric sums, all-reduced over ranks when torch.distributed is initialised
        (the only collective of the path: envs never exchange data)."""
        import torch.distributed as dist
        t = torch.tensor([len(self.rewards), float(sum(self.successes)), float(sum(self.efficiency_scores)),
                          float(sum(self.effort_scores)), float(sum(self.rewards))], dtype=torch.float64,
                         device=self._dev if self._dev is not None else "cpu")
        if dist.is_available() and dist.is_initialized():
            if dist.get_backend() == "gloo":
                t = t.cpu()
            elif not t.is_cuda:
                t = t.cuda()                      # NCCL reduces device tensors (scores kept on the host)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n = max(float(t[0]), 1.0)
        return {"episodes": int(t[0]), "success_rate": float(t[1]) / n, "efficiency_mean": float(t[2]) / n,
                "effort_mean": float(t[3]) / n, "reward_mean": float(t[4]) / n}


class MazeNamoMetric(ShipIceMetric):
    """maze_namo_metric.py:5-75, batched.  Same scores as :class:`ShipIceMetric` with the robot in place of the
    ship, except the reference path length: ``L = goal_dt[int(y0 * s), int(x0 * s)] / s`` read from
    ``info['goal_dt']`` (the un-normalised wavefront hop-count map) at the 2-decimal-rounded start position
    (:69-75)."""

    def __init__(self, alg_name, robot_mass, num_envs=None, device=None) -> None:
        super().__init__(alg_name=alg_name, ship_mass=robot_mass, goal=(0.0, 0.0), num_envs=num_envs, device=device)
        self.robot_mass = float(robot_mass)
        self._goal_dt = None

    def reset(self, info, mask=None):
        super().reset(info, mask)
        if self._goal_dt is None or self._goal_dt_src is not info["goal_dt"]:
            self._goal_dt_src = info["goal_dt"]
            self._goal_dt = torch.as_tensor(info["goal_dt"], dtype=torch.float64, device=self._dev)
        s = float(info["m_to_pix_scale"])
        m = torch.ones(self._n, dtype=torch.bool, device=self._dev) if mask is None else mask.bool()
        st = info["state"]
        px = (st[:, 0] * s).to(torch.int64)       # int() truncates toward zero, like tensor.to(int64)
        py = (st[:, 1] * s).to(torch.int64)
        H, W = self._goal_dt.shape
        px = torch.where(px < 0, px + W, px).clamp(0, W - 1)
        py = torch.where(py < 0, py + H, py).clamp(0, H - 1)
        self.L = torch.where(m, self._goal_dt[py, px] / s, self.L)

    @property
    def total_robot_dist(self):
        return self.total_ship_dist


def _mst_cost(weights) -> float:
    """Total weight of a minimum spanning forest of a small dense graph (Prim; ``weights[i][j] = inf`` = no edge).
    The total is unique whatever the tie-breaking, so it equals ``sum(nx.minimum_spanning_tree(G) weights)``."""
    import numpy as np
    w = np.asarray(weights, dtype=np.float64)
    n = w.shape[0]
    seen = np.zeros(n, dtype=bool)
    total = 0.0
    for root in range(n):
        if seen[root]:
            continue
        seen[root] = True
        best = w[root].copy()
        while True:
            cand = np.where(seen, np.inf, best)
            j = int(np.argmin(cand))
            if not np.isfinite(cand[j]):
                break
            total += float(cand[j])
            seen[j] = True
            best = np.minimum(best, w[j])
    return total


class TaskDrivenMetric(BaseMetric):
    """task_driven_metric.py:8-157 (box-delivery / area-clearing), batched over environments.

    success = completed boxes / boxes; efficiency = MST(robot start, completed boxes' initial centroids, each one's
    nearest goal point) / robot path length (:29-96, the path integrates the 2-dp rounded ``info['state']``);
    effort = (m_0 l_0 + sum_completed d_min * m) / (m_0 l_0 + total_work) (:99-120).
    Reference quirk (D11): ``self.box_mass`` is only assigned when passed but read unconditionally, so the 2D
    policies (which never pass it) hit an AttributeError once a box completes; the intended behaviour —
    ``box_mass=None`` => weigh by polygon area — is what this class does.
    ``info['obs']`` is the batched tensor ``[E, NB, 4, 2]`` with ``info['n_boxes_left']`` valid rows."""

    def __init__(self, alg_name, robot_mass, box_mass=None) -> None:
        super().__init__(alg_name=alg_name)
        self.robot_mass = float(robot_mass)
        self.box_mass = box_mass
        self._n = None

    @staticmethod
    def _np(x):
        import numpy as np
        return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)

    def reset(self, info, mask=None):
        import numpy as np
        st = self._np(info["state"]).astype(np.float64)
        n = st.shape[0]
        if self._n is None:
            self._n = n
            self.eps_reward = np.zeros(n); self.total_mass_dist = np.zeros(n); self.total_robot_dist = np.zeros(n)
            self.robot_state = np.zeros((n, 3)); self.initial_robot_state = np.zeros((n, 3))
            self.centroids = [None] * n; self.areas = [None] * n; self.box_completed_statuses = [None] * n
        m = np.ones(n, dtype=bool) if mask is None else self._np(mask).astype(bool)
        boxes = self._np(info["obs"]).astype(np.float64)
        left = self._np(info["n_boxes_left"])
        self.goal_positions = [[float(g[0]), float(g[1])] for g in info["goal_positions"]]
        for e in np.nonzero(m)[0]:
            self.eps_reward[e] = 0.0; self.total_mass_dist[e] = 0.0; self.total_robot_dist[e] = 0.0
            self.robot_state[e] = st[e]; self.initial_robot_state[e] = st[e]
            polys = boxes[e, :int(left[e])]
            # shapely Polygon(box).area / .centroid
            x, y = polys[:, :, 0], polys[:, :, 1]
            xn, yn = np.roll(x, -1, axis=1), np.roll(y, -1, axis=1)
            cr = x * yn - xn * y
            a2 = cr.sum(axis=1)
            self.areas[e] = np.abs(a2) / 2
            self.centroids[e] = np.stack([((x + xn) * cr).sum(axis=1) / (3 * a2), ((y + yn) * cr).sum(axis=1) / (3 * a2)], axis=1)
            self.box_completed_statuses[e] = np.zeros(len(polys), dtype=bool)

    def compute_mst_cost_for_successful_boxes(self, e):
        import numpy as np
        done = np.nonzero(self.box_completed_statuses[e])[0]
        if len(done) == 0:
            return 0.0
        pos = self.centroids[e][done]
        k = len(done)
        goals = np.asarray(self.goal_positions, dtype=np.float64)
        w = np.full((2 * k + 1, 2 * k + 1), np.inf)
        for i in range(k):
            for j in range(i + 1, k):
                w[i, j] = w[j, i] = np.linalg.norm(pos[i] - pos[j])
            w[2 * k, i] = w[i, 2 * k] = np.linalg.norm(self.initial_robot_state[e, :2] - pos[i])
            w[i, i + k] = w[i + k, i] = np.min(np.linalg.norm(goals - pos[i], axis=1))
        return _mst_cost(w)

    def compute_efficiency_score(self, e, mst_cost):
        s = self.box_completed_statuses[e]
        return float(s.sum()) / len(s), mst_cost / self.total_robot_dist[e]

    def compute_effort_score(self, e):
        import numpy as np
        done = np.nonzero(self.box_completed_statuses[e])[0]
        goals = np.asarray(self.goal_positions, dtype=np.float64)
        min_mass_dist = 0.0
        for i in done:
            d = np.min(np.linalg.norm(goals - self.centroids[e][i], axis=1))
            min_mass_dist += d * (self.box_mass if self.box_mass is not None else self.areas[e][i])
        ml = self.robot_mass * self.total_robot_dist[e]
        return (ml + min_mass_dist) / (ml + self.total_mass_dist[e])

    def update(self, info, reward, eps_complete=False):
        import numpy as np
        self.eps_reward += self._np(reward).astype(np.float64)
        self.total_mass_dist = self._np(info["total_work"]).astype(np.float64).copy()
        comp = self._np(info["box_completed_statuses"]).astype(bool)
        st = self._np(info["state"]).astype(np.float64)
        self.total_robot_dist += np.linalg.norm(self.robot_state[:, :2] - st[:, :2], axis=1)
        self.robot_state = st.copy()
        done = self._np(eps_complete).astype(bool) if not isinstance(eps_complete, bool) else np.full(self._n, eps_complete)
        for e in range(self._n):
            self.box_completed_statuses[e] = comp[e, :len(self.box_completed_statuses[e])]
        for e in np.nonzero(done)[0]:
            self.rewards.append(float(self.eps_reward[e]))
            sr, eff = self.compute_efficiency_score(e, self.compute_mst_cost_for_successful_boxes(e))
            self.success_rates.append(sr)
            self.efficiency_scores.append(float(eff))
            self.effort_scores.append(float(self.compute_effort_score(e)))
        return done

    def reduce(self):
        """Episode-metric sums, all-reduced over ranks when torch.distributed is initialised."""
        import torch.distributed as dist
        t = torch.tensor([len(self.rewards), float(sum(self.success_rates)), float(sum(self.efficiency_scores)),
                          float(sum(self.effort_scores)), float(sum(self.rewards))], dtype=torch.float64)
        if dist.is_available() and dist.is_initialized():
            if dist.get_backend() != "gloo":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n = max(float(t[0]), 1.0)
        return {"episodes": int(t[0]), "success_rate": float(t[1]) / n, "efficiency_mean": float(t[2]) / n,
                "effort_mean": float(t[3]) / n, "reward_mean": float(t[4]) / n}
