from .batch_sampler import BatchSampler, Sampler

__all__ = ["Sampler", "BatchSampler"]
