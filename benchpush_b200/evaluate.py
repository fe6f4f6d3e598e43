"""Batched evaluation loop: what ``evaluate()`` of the reference policies does for one env at a time
(e.g. baselines/box_delivery/ppo/policy.py:73-108: reset, metric.reset(info), step until done or truncated,
metric.update(..., eps_complete) and finally the ``(success_rates, efficiency_scores, effort_scores, rewards)`` lists that
``BaseMetric.plot_algs_scores`` consumes, base_metric.py:107-137) — for a vector env: every env runs its own episode, an
env that finished is scored and restarted while the others keep stepping."""
from __future__ import annotations

import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def evaluate(env, policy, metric, num_episodes: int, max_steps: int = 100000):
    """``policy(obs) -> actions`` (batched; obs is the env's device tensor).  Works with the room envs (explicit
    ``reset(mask=...)``) and with the auto-resetting ship-ice / maze envs.  Returns the metric's score lists."""
    obs, info = env.reset()
    metric.reset(info)
    auto = bool(getattr(env, "auto_reset", False))
    prev_done = None
    for _ in range(max_steps):
        if len(metric.rewards) >= num_episodes:
            break
        obs, rew, term, trunc, info = env.step(policy(obs))
        if auto and prev_done is not None and bool(prev_done.any()):
            # next-step auto reset: for the envs that finished in the previous step this call returned the reset state of their
            # next episode (reward 0).  Start their score accumulators from it; the update below is then a no-op for them.
            metric.reset(info, mask=prev_done)
        done = _np(term).astype(bool) | _np(trunc).astype(bool)
        done_t = (term | trunc) if hasattr(term, "device") else done
        metric.update(info, rew, eps_complete=done_t)
        prev_done = done_t.clone() if hasattr(done_t, "clone") else done.copy()
        if done.any() and not auto:
            obs, info = env.reset(mask=done_t if hasattr(done_t, "device") else done.astype(np.uint8))
            metric.reset(info, mask=done_t)
    n = num_episodes
    return {"success_rates": list(getattr(metric, "success_rates", []) or getattr(metric, "successes", []))[:n],
            "efficiency_scores": metric.efficiency_scores[:n], "effort_scores": metric.effort_scores[:n], "rewards": metric.rewards[:n]}
