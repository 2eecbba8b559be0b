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

    def __add__(self, other):
        """sum of independent Gaussians (per-component posteriors, test.py:403-405); covariances are added densely"""
        if not isinstance(other, MultivariateNormal):
            return MultivariateNormal(self.loc + other, self._covar)
        return MultivariateNormal(self.loc + other.loc, DenseLazyTensor(self._covar.evaluate() + other._covar.evaluate()))

    def sample(self, sample_shape=torch.Size()):
        """draws through a dense Cholesky factor (small n* only), jitter added until the factorisation succeeds"""
        cov = self._covar.evaluate().detach()
        cov = 0.5 * (cov + cov.transpose(-1, -2))
        eye = torch.eye(cov.shape[-1], dtype=cov.dtype, device=cov.device)
        jitter = 1e-6 * float(cov.diagonal().mean().abs().clamp_min(1e-12))
        for _ in range(8):
            try:
                L = torch.linalg.cholesky(cov + jitter * eye)
                break
            except RuntimeError:
                jitter *= 10
        else:
            raise RuntimeError("sample: covariance is not positive definite")
        z = torch.randn(tuple(sample_shape) + (cov.shape[-1],), dtype=cov.dtype, device=cov.device)
        return self.loc + z @ L.transpose(-1, -2)

    def __getitem__(self, idx):
        return MultivariateNormal(self.loc[idx], self._covar.evaluate()[idx][:, idx])
