from torch import Tensor, nn
from torch.optim import Optimizer


class ValueFunction(nn.Module):
    """State-value function: a network plus the optimizer that trains it (ref: value_function.py:5-28)."""

    def __init__(self, network: nn.Module, optimizer: Optimizer) -> None:
        super().__init__()
        self.network = network
        self.optimizer = optimizer

    def forward(self, observation: Tensor) -> Tensor:
        return self.network(observation)
