"""Policy classes with the reference's public surface (ref: policies/*.py).

Action SAMPLING stays on the host through torch CPU ops, exactly like the reference
(ref: policies/stochastic_policy.py:26-41): a GPU Philox stream cannot reproduce torch's CPU mt19937 draws, and
BASELINE.json asks for bit-exact sampled action indices.  The device copy of the parameters is written back into
these host modules after every train().
"""
from abc import ABC, abstractmethod

import numpy as np
import torch
from torch import Tensor, nn
from torch.distributions import Categorical, Distribution, Independent, Normal
from torch.optim import Optimizer


class Policy(nn.Module, ABC):
    """ref: policies/policy.py:8-30"""

    @abstractmethod
    def get_action_tensor(self, observation: Tensor) -> Tensor:
        raise NotImplementedError

    @abstractmethod
    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        raise NotImplementedError


class StochasticPolicy(Policy):
    """ref: policies/stochastic_policy.py:11-41"""

    @abstractmethod
    def forward(self, observation: Tensor) -> Distribution:
        raise NotImplementedError

    def get_action_tensor(self, observation: Tensor) -> Tensor:
        with torch.no_grad():
            dist = self.forward(observation)
        return dist.sample()

    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        action = self.get_action_tensor(torch.from_numpy(observation).float())
        return np.asarray(action.detach().numpy())


class CategoricalPolicy(StochasticPolicy):
    """ref: policies/categorical_policy.py:8-32"""

    def __init__(self, network: nn.Module, optimizer: Optimizer):
        super().__init__()
        self.network = network
        self.optimizer = optimizer

    def forward(self, observation: Tensor) -> Categorical:
        return Categorical(logits=self.network(observation))


class GaussianPolicy(StochasticPolicy):
    """ref: policies/gaussian_policy.py:9-37 -- diagonal Gaussian, mean from the network, std = exp(log_std)."""

    def __init__(self, network: nn.Module, optimizer: Optimizer, log_std: nn.Parameter):
        super().__init__()
        self.network = network
        self.optimizer = optimizer
        self.log_std = log_std

    def forward(self, observation: Tensor) -> Independent:
        return Independent(Normal(loc=self.network(observation), scale=torch.exp(self.log_std)), 1)


class DeterministicPolicy(Policy):
    """ref: policies/deterministic_policy.py:9-45"""

    def __init__(self, network: nn.Module, optimizer: Optimizer):
        super().__init__()
        self.network = network
        self.optimizer = optimizer

    def forward(self, observation: Tensor) -> Tensor:
        return self.network(observation)

    def get_action_tensor(self, observation: Tensor) -> Tensor:
        with torch.no_grad():
            return self.forward(observation)

    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        return np.asarray(self.get_action_tensor(torch.from_numpy(observation).float()).detach().numpy())


class RandomPolicy(Policy):
    """Uniform samples from the action space, for warm-up (ref: policies/random_policy.py:9-27)."""

    def __init__(self, action_space) -> None:
        super().__init__()
        self.action_space = action_space

    def get_action_tensor(self, observation: Tensor) -> Tensor:
        return torch.from_numpy(self.action_space.sample())

    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        return np.asarray(self.action_space.sample())
