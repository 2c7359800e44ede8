from .base import (CategoricalPolicy, DeterministicPolicy, GaussianPolicy, Policy, RandomPolicy, StochasticPolicy)

__all__ = ["Policy", "StochasticPolicy", "CategoricalPolicy", "GaussianPolicy", "DeterministicPolicy", "RandomPolicy"]
