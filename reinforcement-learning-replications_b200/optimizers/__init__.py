from .conjugate_gradient_optimizer import ConjugateGradientOptimizer

__all__ = ["ConjugateGradientOptimizer"]
