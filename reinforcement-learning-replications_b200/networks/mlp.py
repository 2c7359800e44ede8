"""Fully connected network in the shape the update engine consumes: Linear layers with one activation type between
them and another after the last (ref: networks/mlp.py:15-41).  The module tree -- ``network`` = Sequential with the
Linear layers at the even indices -- is part of the checkpoint format (``network.0.weight`` ...), so it is kept."""
from typing import Iterator, List, Sequence, Tuple, Type

from torch import Tensor, nn


def _layer_dims(sizes: Sequence[int]) -> Iterator[Tuple[int, int]]:
    widths = [int(w) for w in sizes]
    if len(widths) < 2:
        raise ValueError("MLP needs at least an input and an output width")
    return zip(widths, widths[1:])


class MLP(nn.Module):
    def __init__(self, sizes: Sequence[int], activation_function: Type[nn.Module] = nn.Tanh,
                 output_activation_function: Type[nn.Module] = nn.Identity) -> None:
        super().__init__()
        self.sizes: List[int] = [int(w) for w in sizes]
        stack: List[nn.Module] = []
        for fan_in, fan_out in _layer_dims(self.sizes):
            stack += [nn.Linear(fan_in, fan_out), activation_function()]
        stack[-1] = output_activation_function()  # the activation after the last Linear
        self.network: nn.Module = nn.Sequential(*stack)

    def forward(self, input: Tensor) -> Tensor:
        return self.network(input)
