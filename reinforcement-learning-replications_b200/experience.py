from typing import List, Optional

import numpy as np


class Experience:
    """Trajectory container with the reference's nested-list layout (ref: experience.py:6-84):
    ``observations[e][t]``, ``actions[e][t]``, ``rewards[e][t]``, ``dones[e][t]`` and one ``last_observations[e]``."""

    _FIELDS = ("observations", "actions", "rewards", "last_observations", "dones", "episode_returns", "episode_lengths")

    def __init__(self, observations: Optional[List[List[np.ndarray]]] = None,
                 actions: Optional[List[List[np.ndarray]]] = None, rewards: Optional[List[List[float]]] = None,
                 last_observations: Optional[List[np.ndarray]] = None, dones: Optional[List[List[bool]]] = None,
                 episode_returns: Optional[List[float]] = None, episode_lengths: Optional[List[int]] = None):
        given = (observations, actions, rewards, last_observations, dones, episode_returns, episode_lengths)
        for name, value in zip(self._FIELDS, given):
            setattr(self, name, value if value else [])

    @property
    def observations_with_last_observation(self):
        return [list(obs) + [last] for obs, last in zip(self.observations, self.last_observations)]

    @property
    def next_observations(self):
        return [list(obs[1:]) + [last] for obs, last in zip(self.observations, self.last_observations)]

    @property
    def episode_dones(self) -> List[bool]:
        return [episode[-1] for episode in self.dones]

    @staticmethod
    def _flat(nested):
        return [item for episode in nested for item in episode]

    @property
    def flattened_observations(self):
        return self._flat(self.observations)

    @property
    def flattened_actions(self):
        return self._flat(self.actions)

    @property
    def flattened_rewards(self):
        return self._flat(self.rewards)

    @property
    def flattened_next_observations(self):
        return self._flat(self.next_observations)

    @property
    def flattened_dones(self):
        return self._flat(self.dones)
