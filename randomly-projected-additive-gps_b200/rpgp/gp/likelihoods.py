"""GaussianLikelihood: homoskedastic noise, value = softplus(raw) + 1e-4 (training_routines.py:345-351)."""
import torch

from .constraints import GreaterThan
from .distributions import MultivariateNormal
from .module import Module


class GaussianLikelihood(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, **kwargs):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(1)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))
        if noise_prior is not None:
            self.register_prior("noise_prior", noise_prior, lambda m: m.noise)
        # gpytorch nests the parameter under `noise_covar`; keep that path reachable for state-dict helpers
        self.noise_covar = _NoiseAlias(self)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        self._set_constrained("raw_noise", value)

    def forward(self, dist, *args, **kwargs):
        return MultivariateNormal(dist.mean, dist.lazy_covariance_matrix.evaluate_kernel().add_diag(self.noise))

    __call__ = forward

    def train(self, mode=True):
        return super().train(mode)


class _NoiseAlias:
    """`likelihood.noise_covar.noise` / `.raw_noise` as in gpytorch (not a Module: it must not duplicate parameters)"""

    def __init__(self, owner):
        object.__setattr__(self, "_owner", owner)

    @property
    def noise(self):
        return self._owner.noise

    @property
    def raw_noise(self):
        return self._owner.raw_noise
