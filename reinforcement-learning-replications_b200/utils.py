"""Host-side helpers with the reference's names (ref: utils.py).  The arithmetic-heavy ones of the update path
(discounted_cumulative_sums, gae, compute_values, normalize_tensor, polyak_average inside train()) run inside the
native engine; the functions here are the host-side conveniences the class API needs."""
import random

import numpy as np
import torch
from torch import Tensor

from .policies import Policy


class _NoisedPolicy(Policy):
    """Gaussian exploration noise around a deterministic policy, clipped to the action limit
    (ref: utils.py:101-124; numpy global RNG for get_action_numpy, torch RNG for get_action_tensor)."""

    def __init__(self, base_policy: Policy, action_space, action_noise_scale: float):
        super().__init__()
        self.base_policy = base_policy
        self.action_space = action_space
        self.action_noise_scale = action_noise_scale
        self.action_limit = action_space.high[0]
        self.action_size = action_space.shape[0]

    def get_action_tensor(self, observation: Tensor) -> Tensor:
        action = self.base_policy.get_action_tensor(observation)
        action += self.action_noise_scale * torch.randn(self.action_size)
        return torch.clip(action, -self.action_limit, self.action_limit)

    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        action = self.base_policy.get_action_numpy(observation)
        action += self.action_noise_scale * np.random.randn(self.action_size)
        return np.clip(action, -self.action_limit, self.action_limit)


def add_noise_to_get_action(policy: Policy, action_space, action_noise_scale: float) -> Policy:
    return _NoisedPolicy(policy, action_space, action_noise_scale)


def set_seed_for_libraries(seed: int) -> None:
    """Seed python / numpy / torch and make torch deterministic (ref: utils.py:127-136)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.use_deterministic_algorithms(True)
