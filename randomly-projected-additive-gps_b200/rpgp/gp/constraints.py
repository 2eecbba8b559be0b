"""Parameter constraints and priors (GPyTorch semantics: Positive = softplus; raw parameters initialised at 0)."""
import math

import torch
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Positive(torch.nn.Module):
    lower_bound = 0.0

    def transform(self, raw):
        return softplus(raw)

    def inverse_transform(self, value):
        return inv_softplus(value)


class GreaterThan(Positive):
    """value = softplus(raw) + lower_bound (GaussianLikelihood noise uses GreaterThan(1e-4))"""

    def __init__(self, lower_bound):
        super().__init__()
        self.lower_bound = float(lower_bound)

    def transform(self, raw):
        return softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return inv_softplus(value - self.lower_bound)


class SmoothedBoxPrior(torch.nn.Module):
    """gpytorch.priors.SmoothedBoxPrior(a, b, sigma): flat on [a, b] with Gaussian shoulders
    (training_routines.py:345-350 puts it on the noise)."""

    def __init__(self, a, b, sigma=0.01):
        super().__init__()
        self.a, self.b, self.sigma = float(a), float(b), float(sigma)

    def log_prob(self, x):
        c = 0.5 * (self.a + self.b)
        r = 0.5 * (self.b - self.a)
        m = 1.0 + (self.b - self.a) / (math.sqrt(2.0 * math.pi) * self.sigma)
        xt = ((x - c).abs() - r).clamp(min=0)
        normal_lp = -0.5 * (xt / self.sigma) ** 2 - math.log(self.sigma) - 0.5 * math.log(2.0 * math.pi)
        return normal_lp - math.log(m)
