"""ConjugateGradientOptimizer with the reference's constructor and (hyper-parameter only) state
(ref: optimizers/conjugate_gradient_optimizer.py:24-131).

In the reference, ``step(loss_fn, kl_fn)`` evaluates two torch closures many times (double backprop for every
Hessian-vector product).  Here the whole step -- surrogate gradient, 11 analytic Fisher-vector products, conjugate
gradient, step size, backtracking line search with reject/restore -- runs natively inside ``TRPO.train``
(b200rl_trpo_update); this class carries the hyper-parameters and the parameter list for it.
"""
import logging
from typing import Iterable

from torch import Tensor
from torch.optim import Optimizer

logger = logging.getLogger(__name__)

_DEFAULTS = dict(max_constraint=0.01, n_conjugate_gradients=10, max_backtracks=15, backtrack_ratio=0.8,
                 hvp_damping_coefficient=1e-5)


class ConjugateGradientOptimizer(Optimizer):
    def __init__(self, params: Iterable[Tensor], max_constraint: float = 0.01, n_conjugate_gradients: int = 10,
                 max_backtracks: int = 15, backtrack_ratio: float = 0.8, hvp_damping_coefficient: float = 1e-5):
        super().__init__(params, {})
        self.max_constraint = max_constraint
        self.n_conjugate_gradients = n_conjugate_gradients
        self.max_backtracks = max_backtracks
        self.backtrack_ratio = backtrack_ratio
        self.hvp_damping_coefficient = hvp_damping_coefficient

    def hyper_parameters(self) -> dict:
        return {k: getattr(self, k) for k in _DEFAULTS}

    @property
    def state(self) -> dict:  # only hyper-parameters are serialised, like the reference (:100-119)
        return self.hyper_parameters()

    @state.setter
    def state(self, state: dict) -> None:
        for k, default in _DEFAULTS.items():
            setattr(self, k, state.get(k, default))

    def __setstate__(self, state: dict) -> None:
        if "hvp_damping_coefficient" not in state["state"]:
            logger.warning("Resuming ConjugateGradientOptimizer with lost state.")
        self.state = state["state"]
        self.param_groups = state["param_groups"]

    def step(self, loss_function=None, kl_divergence_function=None) -> None:
        raise NotImplementedError(
            "The constrained step runs natively inside rl_replicas_b200.algorithms.TRPO.train(); arbitrary torch "
            "closures cannot be executed by the B200 engine.")
