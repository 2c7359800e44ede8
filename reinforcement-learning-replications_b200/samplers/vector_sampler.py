"""Vectorised host sampler: the caller of the update path when a batch is 1024 envs x 1000 steps.

``BatchSampler`` (ref: samplers/batch_sampler.py:31-101) steps ONE environment and calls the policy once per step; at
the benchmark shapes that is a million tiny torch calls per epoch.  ``VectorSampler`` keeps the same
``Sampler.sample(num_samples, policy) -> Experience`` interface, steps ``len(envs)`` environments in lock step and asks
the policy for all their actions in ONE batched call per step (``get_action_numpy`` on an ``[E, O]`` array -- the
reference's policies are batch-transparent: the network and the distribution broadcast over leading dimensions).
Rollouts go straight into one ``PackedExperience`` (the engine's contiguous layout), environment by environment.

Semantics kept from the reference: an episode closes on terminated OR truncated (:65) or when the environment's share
of the epoch is used up (cut-off: ``done`` stays False and the next observation is the episode's ``last_observation``);
``is_continuous`` carries the running observation across ``sample`` calls, otherwise every call starts with a reset.
Random streams differ from E independent ``BatchSampler``s by construction (one batched draw per step instead of E
scalar draws), so parity with the reference is distributional, not bit-wise -- which is why this is opt-in.
"""
from typing import List, Optional, Sequence

import numpy as np

from ..experience import Experience, PackedExperience
from .batch_sampler import Sampler


class VectorSampler(Sampler):
    def __init__(self, envs: Sequence, seed: Optional[int] = None, is_continuous: bool = False, pinned: bool = False):
        if len(envs) == 0:
            raise ValueError("VectorSampler needs at least one environment")
        self.envs = list(envs)
        self.seed = seed
        self.is_continuous = is_continuous
        self.pinned = pinned
        self.observations: Optional[List[np.ndarray]] = None

    def sample(self, num_samples: int, policy) -> Experience:
        """``num_samples`` transitions in total, split as evenly as possible over the environments."""
        n_env = len(self.envs)
        steps = [num_samples // n_env + (1 if e < num_samples % n_env else 0) for e in range(n_env)]
        if self.observations is None:
            self.observations = [env.reset(seed=None if self.seed is None else self.seed + e)[0]
                                 for e, env in enumerate(self.envs)]
        elif not self.is_continuous:
            self.observations = [env.reset()[0] for env in self.envs]
        obs_dim = int(np.asarray(self.observations[0]).size)
        # per-environment scratch in [E, T, .] arrays, flushed environment by environment at the end so that every
        # episode is contiguous in the packed store
        horizon = max(steps)
        probe = np.asarray(policy.get_action_numpy(np.stack(self.observations).astype(np.float32)))
        act_dim = int(probe[0].size)
        obs_buf = np.empty((n_env, horizon, obs_dim), np.float32)
        act_buf = np.empty((n_env, horizon, max(act_dim, 1)), np.float32)
        rew_buf = np.empty((n_env, horizon), np.float64)
        done_buf = np.zeros((n_env, horizon), bool)
        last_obs: List[List[np.ndarray]] = [[] for _ in range(n_env)]  # one per closed episode
        ends: List[List[int]] = [[] for _ in range(n_env)]
        actions = probe
        for t in range(horizon):
            if t > 0:
                actions = np.asarray(policy.get_action_numpy(np.stack(self.observations).astype(np.float32)))
            for e, env in enumerate(self.envs):
                if t >= steps[e]:
                    continue
                obs_buf[e, t] = np.asarray(self.observations[e], np.float32).reshape(-1)
                act_buf[e, t] = np.asarray(actions[e], np.float32).reshape(-1)
                self.observations[e], reward, terminated, truncated, _ = env.step(actions[e])
                finished = bool(terminated or truncated)
                rew_buf[e, t] = reward
                done_buf[e, t] = finished
                if finished or t == steps[e] - 1:
                    ends[e].append(t + 1)
                    last_obs[e].append(np.asarray(self.observations[e], np.float32).reshape(-1).copy())
                    if finished:
                        self.observations[e], _ = env.reset()
        exp = PackedExperience(num_samples, obs_dim, act_dim, self.pinned, scalar_actions=np.asarray(probe[0]).ndim == 0)
        for e in range(n_env):  # one block copy per episode (ref batch_sampler.py:66-78 closes episodes one by one)
            begin = 0
            for end, last in zip(ends[e], last_obs[e]):
                exp.append_episode(obs_buf[e, begin:end], act_buf[e, begin:end], rew_buf[e, begin:end],
                                   done_buf[e, begin:end], last)
                begin = end
        return exp
