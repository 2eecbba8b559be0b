"""Mean functions (gpytorch.means.ConstantMean; gp_models/models.py:14)."""
import torch

from .module import Module


class ConstantMean(Module):
    def __init__(self, prior=None):
        super().__init__()
        self.register_parameter("constant", torch.nn.Parameter(torch.zeros(1)))
        if prior is not None:
            self.register_prior("mean_prior", prior, lambda m: m.constant)

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class ZeroMean(Module):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)
