"""Synthetic trajectory batches shaped like BASELINE.json's configs.

There is no gymnasium / MuJoCo in the image, so every benchmark and parity input is
synthetic: fixed-seed numpy data in the *packed* layout the engine consumes
(obs [N,O] f32, act [N,A] | [N] f32, rew [N] f64, last_obs [E,O] f32,
ep_offsets [E+1] i64, ep_done [E] bool).  ``to_experience`` converts a packed batch to
the reference's nested-list ``Experience`` layout (ref: experience.py:6-40).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def fixed_batch(n_envs: int, horizon: int, obs_dim: int, act_dim: int, discrete: bool = False, seed: int = 0,
                frac_not_done: float = 0.0, mean_fn=None, log_std: float = -0.5) -> Dict[str, np.ndarray]:
    """n_envs episodes of exactly ``horizon`` steps (BASELINE configs 2/3/5)."""
    rng = np.random.default_rng(seed)
    n = n_envs * horizon
    obs = rng.standard_normal((n, obs_dim), dtype=np.float32)
    if discrete:
        act = rng.integers(0, act_dim, size=n).astype(np.float32)
    else:
        noise = rng.standard_normal((n, act_dim), dtype=np.float32) * np.float32(np.exp(log_std))
        act = (mean_fn(obs) + noise).astype(np.float32) if mean_fn is not None else noise
    rew = rng.standard_normal(n)  # float64, like Python-float rewards
    last_obs = rng.standard_normal((n_envs, obs_dim), dtype=np.float32)
    done = np.ones(n_envs, dtype=bool)
    if frac_not_done > 0:
        done[rng.random(n_envs) < frac_not_done] = False
    offsets = (np.arange(n_envs + 1, dtype=np.int64) * horizon)
    return dict(obs=obs, act=act, rew=rew, last_obs=last_obs, ep_offsets=offsets, ep_done=done)


def ragged_batch(n_total: int, obs_dim: int, act_dim: int, discrete: bool, seed: int = 0, min_len: int = 10,
                 max_len: int = 200, mean_fn=None, log_std: float = -0.5, reward_const: Optional[float] = None
                 ) -> Dict[str, np.ndarray]:
    """Variable-length episodes summing to n_total; the last one is cut and not done (BASELINE config 1)."""
    rng = np.random.default_rng(seed)
    lens = []
    while sum(lens) < n_total:
        lens.append(int(rng.integers(min_len, max_len)))
    lens[-1] -= sum(lens) - n_total
    if lens[-1] == 0:
        lens.pop()
    e = len(lens)
    obs = rng.standard_normal((n_total, obs_dim), dtype=np.float32)
    if discrete:
        act = rng.integers(0, act_dim, size=n_total).astype(np.float32)
    else:
        noise = rng.standard_normal((n_total, act_dim), dtype=np.float32) * np.float32(np.exp(log_std))
        act = (mean_fn(obs) + noise).astype(np.float32) if mean_fn is not None else noise
    rew = np.full(n_total, reward_const, dtype=np.float64) if reward_const is not None else rng.standard_normal(n_total)
    last_obs = rng.standard_normal((e, obs_dim), dtype=np.float32)
    done = np.ones(e, dtype=bool)
    done[-1] = False
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return dict(obs=obs, act=act, rew=rew, last_obs=last_obs, ep_offsets=offsets, ep_done=done)


def shard_batch(batch: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    """Contiguous block of episodes for ``rank`` (shard by environment)."""
    e = len(batch["ep_done"])
    lo, hi = (e * rank) // world, (e * (rank + 1)) // world
    off = batch["ep_offsets"]
    s, t = int(off[lo]), int(off[hi])
    return dict(obs=batch["obs"][s:t], act=batch["act"][s:t], rew=batch["rew"][s:t], last_obs=batch["last_obs"][lo:hi],
                ep_offsets=(off[lo:hi + 1] - off[lo]).astype(np.int64), ep_done=batch["ep_done"][lo:hi])


def to_experience_lists(batch: Dict[str, np.ndarray], discrete: bool):
    """Packed batch -> the nested Python lists of the reference's Experience (ref: experience.py:18-40)."""
    off = batch["ep_offsets"]
    observations, actions, rewards, dones = [], [], [], []
    for e in range(len(off) - 1):
        s, t = int(off[e]), int(off[e + 1])
        observations.append([batch["obs"][i] for i in range(s, t)])
        if discrete:
            actions.append([np.int64(batch["act"][i]) for i in range(s, t)])
        else:
            actions.append([batch["act"][i] for i in range(s, t)])
        rewards.append([float(x) for x in batch["rew"][s:t]])
        d = [False] * (t - s)
        d[-1] = bool(batch["ep_done"][e])
        dones.append(d)
    last_observations = [batch["last_obs"][e] for e in range(len(off) - 1)]
    return dict(observations=observations, actions=actions, rewards=rewards, last_observations=last_observations,
                dones=dones, episode_returns=[float(sum(r)) for r in rewards],
                episode_lengths=[len(r) for r in rewards])
