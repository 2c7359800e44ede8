from .batch_sampler import BatchSampler, Sampler
from .vector_sampler import VectorSampler

__all__ = ["Sampler", "BatchSampler", "VectorSampler"]
