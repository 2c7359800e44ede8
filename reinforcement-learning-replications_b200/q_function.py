import torch
from torch import Tensor, nn
from torch.optim import Optimizer


class QFunction(nn.Module):
    """Action-value function Q(s, a) = net([s, a]) with the trailing unit axis removed (ref: q_function.py:6-32)."""

    def __init__(self, network: nn.Module, optimizer: Optimizer) -> None:
        super().__init__()
        self.network = network
        self.optimizer = optimizer

    def forward(self, observation: Tensor, action: Tensor) -> Tensor:
        return self.network(torch.cat((observation, action), dim=-1)).squeeze(-1)
