"""rpgp.gp -- the slice of GPyTorch's module API that the reference's kernel / model / training code touches,
re-implemented on top of the fused K.V operator (GPyTorch itself is not a dependency).

Written so that reference-style code reads unchanged after `from rpgp import gp as gpytorch`:
    gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel()), gpytorch.likelihoods.GaussianLikelihood(),
    gpytorch.mlls.ExactMarginalLogLikelihood(likelihood, model), gpytorch.settings.cg_tolerance(0.002), ...
"""
from . import constraints, distributions, kernels, likelihoods, means, mlls, models, settings  # noqa: F401
from .constraints import SmoothedBoxPrior
from .module import Module  # noqa: F401


class priors:  # namespace shim: gpytorch.priors.SmoothedBoxPrior
    SmoothedBoxPrior = SmoothedBoxPrior
