"""MultivariateNormal with a lazy covariance (gpytorch.distributions.MultivariateNormal; gp_models/models.py:20)."""
import math

import torch

from ..lazy import DenseLazyTensor, LazyTensor


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.loc = mean
        self._covar = covariance_matrix if isinstance(covariance_matrix, LazyTensor) else DenseLazyTensor(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def variance(self):
        return self._covar.diag().clamp_min(1e-10)  # gpytorch clamps negative/zero variances to min_variance

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        """mean -/+ 2 standard deviations (training_routines.py:572-573)"""
        std2 = self.stddev * 2
        return self.mean - std2, self.mean + std2

    def log_prob(self, value):
        diff = value - self.loc
        covar = self._covar.evaluate_kernel()
        if not hasattr(covar, "inv_quad_logdet"):
            covar = covar.add_diag(torch.zeros((), dtype=diff.dtype, device=diff.device))
        inv_quad, logdet = covar.inv_quad_logdet(inv_quad_rhs=diff.unsqueeze(-1), logdet=True)
        return -0.5 * (inv_quad + logdet + diff.shape[-1] * math.log(2 * math.pi))

    def __getitem__(self, idx):
        return MultivariateNormal(self.loc[idx], self._covar.evaluate()[idx][:, idx])
