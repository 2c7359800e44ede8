"""Host-side rollout collection (ref: samplers/sampler.py:7-21, samplers/batch_sampler.py:12-101).

Environment stepping stays on the host (BASELINE.json north_star); this is the producer of the update path's input
and is duck-typed on the gymnasium Env protocol (reset(seed=) -> (obs, info); step(a) -> 5-tuple) so that it needs
no gymnasium import.
"""
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np

from ..experience import Experience, PackedExperience


class Sampler(ABC):
    @abstractmethod
    def sample(self, num_samples: int, policy) -> Experience:
        raise NotImplementedError


class _Episode:
    """Per-episode scratch lists; flushed into the Experience when the episode (or the epoch) ends."""

    def __init__(self):
        self.obs, self.act, self.rew, self.done = [], [], [], []
        self.ret = 0.0

    def flush(self, exp: Experience, last_observation: np.ndarray) -> "_Episode":
        exp.observations.append(self.obs)
        exp.actions.append(self.act)
        exp.rewards.append(self.rew)
        exp.dones.append(self.done)
        exp.last_observations.append(last_observation)
        exp.episode_returns.append(self.ret)
        exp.episode_lengths.append(len(self.rew))
        return _Episode()


class BatchSampler(Sampler):
    """One env, ``num_samples`` steps per call.  The first call seeds the env; later calls reset it unless
    ``is_continuous`` keeps the running observation (ref: batch_sampler.py:49-53).  An episode closes on
    terminated OR truncated (ref: :65) or when the epoch's last step is reached (cut-off, ``done`` stays False)."""

    def __init__(self, env, seed: Optional[int] = None, is_continuous: bool = False, packed: bool = False,
                 pinned: bool = False):
        self.env = env
        self.seed = seed
        self.is_continuous = is_continuous
        self.packed = packed  # write the rollout straight into the engine's contiguous layout (PackedExperience)
        self.pinned = pinned
        self.observation: Optional[np.ndarray] = None

    def _sample_packed(self, num_samples: int, policy) -> Experience:
        """Same control flow as ``sample`` below, writing into a PackedExperience."""
        exp: Optional[PackedExperience] = None
        for step in range(num_samples):
            observation = self.observation
            action = policy.get_action_numpy(observation)
            if exp is None:
                exp = PackedExperience(num_samples, np.asarray(observation).size, np.asarray(action).size, self.pinned,
                                       scalar_actions=np.asarray(action).ndim == 0)
            self.observation, reward, terminated, truncated, _ = self.env.step(action)
            finished = terminated or truncated
            exp.append_step(observation, action, reward, finished)
            if finished or step == num_samples - 1:
                exp.end_episode(self.observation)
                if finished:
                    self.observation, _ = self.env.reset()
        return exp if exp is not None else Experience()

    def sample(self, num_samples: int, policy) -> Experience:
        exp = Experience()
        if self.observation is None:
            self.observation, _ = self.env.reset(seed=self.seed)
        elif not self.is_continuous:
            self.observation, _ = self.env.reset()
        if self.packed:
            return self._sample_packed(num_samples, policy)
        ep = _Episode()
        for step in range(num_samples):
            assert self.observation is not None
            ep.obs.append(self.observation)
            action = policy.get_action_numpy(self.observation)
            ep.act.append(action)
            self.observation, reward, terminated, truncated, _ = self.env.step(action)
            finished = terminated or truncated
            ep.ret += reward
            ep.rew.append(reward)
            ep.done.append(finished)
            if finished or step == num_samples - 1:
                ep = ep.flush(exp, self.observation)
                if finished:
                    self.observation, _ = self.env.reset()
        return exp
