from typing import List, Sequence, Type

from torch import Tensor, nn


class MLP(nn.Module):
    """Multilayer perceptron with the reference's constructor and state_dict layout
    (ref: networks/mlp.py:15-41): ``network`` is a Sequential of Linear/activation pairs, hidden activation
    ``activation_function`` (default Tanh), last activation ``output_activation_function`` (default Identity)."""

    def __init__(self, sizes: Sequence[int], activation_function: Type[nn.Module] = nn.Tanh,
                 output_activation_function: Type[nn.Module] = nn.Identity) -> None:
        super().__init__()
        self.sizes: List[int] = [int(s) for s in sizes]
        modules: List[nn.Module] = []
        last = len(self.sizes) - 2
        for index, (fan_in, fan_out) in enumerate(zip(self.sizes[:-1], self.sizes[1:])):
            modules.append(nn.Linear(fan_in, fan_out))
            modules.append((output_activation_function if index == last else activation_function)())
        self.network: nn.Module = nn.Sequential(*modules)

    def forward(self, input: Tensor) -> Tensor:
        return self.network(input)
