"""Small helpers shared by the kernel factories (reference: gp_models/kernels/etc.py:6-7).  The reference's `DNN`
projection module (deep_rp spec) is a non-linear projection and outside the K.V hot path."""
import torch


def _sample_from_range(num_samples, range_):
    """num_samples uniform draws from [range_[0], range_[1]] (torch global RNG, like the reference)."""
    lo, hi = range_[0], range_[1]
    return torch.rand(num_samples) * (hi - lo) + lo
