from typing import Dict, List

import numpy as np

from .experience import Experience


class ReplayBuffer:
    """FIFO transition store with the reference's interface (ref: replay_buffer.py:9-74): ``add_experience`` appends
    the flattened transitions and drops the oldest beyond ``buffer_size``; ``sample_minibatch`` draws indices with
    ``np.random.randint`` (global numpy RNG, with replacement) -- the same random stream as the reference.
    Storage is Python lists like the reference (the per-call gather is numpy); a device-resident ring buffer is
    SURVEY.md section 8f-4."""

    def __init__(self, buffer_size: int = int(1e6)) -> None:
        self.buffer_size = buffer_size
        self.current_size: int = 0
        self.observations: List[np.ndarray] = []
        self.actions: List[np.ndarray] = []
        self.rewards: List[float] = []
        self.next_observations: List[np.ndarray] = []
        self.dones: List[bool] = []

    def _columns(self):
        return (self.observations, self.actions, self.rewards, self.next_observations, self.dones)

    def add_experience(self, experience: Experience) -> None:
        new = (experience.flattened_observations, experience.flattened_actions, experience.flattened_rewards,
               experience.flattened_next_observations, experience.flattened_dones)
        for column, values in zip(self._columns(), new):
            column.extend(values)
        self.current_size += len(new[0])
        overflow = self.current_size - self.buffer_size
        if overflow > 0:
            for column in self._columns():
                del column[:overflow]
            self.current_size -= overflow

    def sample_minibatch(self, minibatch_size: int = 32) -> Dict[str, np.ndarray]:
        indices = np.random.randint(0, self.current_size, minibatch_size)
        take = lambda column: [column[i] for i in indices]
        return {
            "observations": np.vstack(take(self.observations)),
            "actions": np.vstack(take(self.actions)),
            "rewards": np.asarray(take(self.rewards)),
            "next_observations": np.vstack(take(self.next_observations)),
            "dones": np.asarray(take(self.dones)),
        }
