"""Host-side packing of the reference's nested-list Experience into contiguous arrays.

The reference flattens inside train() with np.concatenate / np.stack per episode
(ref: algorithms/ppo.py:153-154, utils.py:66-70); here that happens once, producing the packed
layout the engine consumes (obs [N,O] f32, act f32, rew f64, last_obs [E,O] f32, CSR offsets, done flags).
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def pack_experience(experience) -> Dict[str, np.ndarray]:
    if hasattr(experience, "packed"):  # PackedExperience: already in the engine's layout, zero copy
        return experience.packed()
    lengths = np.asarray([len(r) for r in experience.rewards], dtype=np.int64)
    if lengths.size == 0 or np.any(lengths == 0):
        raise ValueError("experience must hold at least one episode and no empty episode")
    offsets = np.zeros(lengths.size + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    obs = np.concatenate([np.asarray(o) for o in experience.observations]).astype(np.float32, copy=False)
    if obs.ndim == 1:
        obs = obs[:, None]
    obs = obs.reshape(obs.shape[0], -1)
    act = np.concatenate([np.asarray(a) for a in experience.actions]).astype(np.float32)  # .float(), ppo.py:154
    rew = np.concatenate([np.asarray(r, dtype=np.float64) for r in experience.rewards])
    last = np.stack([np.asarray(o) for o in experience.last_observations]).astype(np.float32, copy=False)
    last = last.reshape(last.shape[0], -1)
    done = np.asarray(experience.episode_dones, dtype=bool)
    return dict(obs=np.ascontiguousarray(obs), act=np.ascontiguousarray(act), rew=rew,
                last_obs=np.ascontiguousarray(last), ep_offsets=offsets, ep_done=done)
