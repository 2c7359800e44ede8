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


# ---- learners on given parameters (bench.py, tools/): plain (W [out, in], b [out]) numpy pairs per layer ------------
def flatten_layers(layers) -> np.ndarray:
    """[(W, b), ...] -> the flat f32 parameter vector in torch's ``parameters()`` order (W row-major, then b)."""
    return np.concatenate([np.concatenate([np.asarray(w, np.float32).reshape(-1), np.asarray(b, np.float32).reshape(-1)])
                           for w, b in layers])


def numpy_mlp(layers, x: np.ndarray, hidden: str = "tanh") -> np.ndarray:
    """Forward pass on the host, used to draw on-policy-like synthetic actions around the initial policy's mean."""
    h = np.asarray(x, np.float32)
    for i, (w, b) in enumerate(layers):
        h = h @ np.asarray(w, np.float32).T + np.asarray(b, np.float32)
        if i < len(layers) - 1:
            h = np.tanh(h) if hidden == "tanh" else np.maximum(h, 0.0)
    return h


def _mlp_from_layers(layers, hidden_activation, output_activation):
    import torch
    from .networks import MLP
    sizes = [int(np.asarray(layers[0][0]).shape[1])] + [int(np.asarray(w).shape[0]) for w, _ in layers]
    net = MLP(sizes, hidden_activation, output_activation)
    linears = [m for m in net.network if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for lin, (w, b) in zip(linears, layers):
            lin.weight.copy_(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)))
            lin.bias.copy_(torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)))
    return net


def onpolicy_learner(algo: str, policy_layers, value_layers, log_std: Optional[np.ndarray] = None, **hparams):
    """PPO / VPG / TRPO on the given initial parameters (Gaussian policy when ``log_std`` is given, else categorical),
    with the recipes' optimizers (Adam 3e-4 / 1e-3, conjugate gradient for TRPO), no env, no sampler, no logging."""
    import torch
    from .algorithms import PPO, TRPO, VPG
    from .optimizers import ConjugateGradientOptimizer
    from .policies import CategoricalPolicy, GaussianPolicy
    from .value_function import ValueFunction
    pnet = _mlp_from_layers(policy_layers, torch.nn.Tanh, torch.nn.Identity)
    vnet = _mlp_from_layers(value_layers, torch.nn.Tanh, torch.nn.Identity)
    popt = (ConjugateGradientOptimizer(pnet.parameters()) if algo == "trpo"
            else torch.optim.Adam(pnet.parameters(), lr=3e-4))
    if log_std is not None:
        policy = GaussianPolicy(pnet, popt, torch.nn.Parameter(torch.from_numpy(np.asarray(log_std, dtype=np.float32))))
    else:
        policy = CategoricalPolicy(pnet, popt)
    value_function = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    learner = {"ppo": PPO, "vpg": VPG, "trpo": TRPO}[algo](policy, value_function, None, None, **hparams)
    learner.metrics_manager = None
    learner.current_total_steps = 0
    return learner


def offpolicy_learner(twin: bool, policy_layers, q_layers_list):
    """TD3 (``twin``) / DDPG on the given initial parameters with an empty ReplayBuffer; returns (learner, buffer)."""
    import types
    import torch
    from .algorithms import DDPG, TD3
    from .policies import DeterministicPolicy, RandomPolicy
    from .q_function import QFunction
    from .replay_buffer import ReplayBuffer
    pnet = _mlp_from_layers(policy_layers, torch.nn.ReLU, torch.nn.Tanh)
    act_dim = int(np.asarray(policy_layers[-1][0]).shape[0])
    policy = DeterministicPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=1e-3))
    critics = []
    for layers in q_layers_list:
        qnet = _mlp_from_layers(layers, torch.nn.ReLU, torch.nn.Identity)
        critics.append(QFunction(qnet, torch.optim.Adam(qnet.parameters(), lr=1e-3)))
    env = types.SimpleNamespace(action_space=types.SimpleNamespace(high=np.ones(act_dim, np.float32), shape=(act_dim,)),
                                spec=types.SimpleNamespace(id="synthetic"))
    buffer = ReplayBuffer()
    if twin:
        learner = TD3(policy, RandomPolicy(None), critics[0], critics[1], env, None, buffer, None)
    else:
        learner = DDPG(policy, RandomPolicy(None), critics[0], env, None, buffer, None)
    learner.metrics_manager = None
    learner.current_total_steps = 0
    return learner, buffer
