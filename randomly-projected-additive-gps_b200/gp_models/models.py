"""Exact GP model with a constant mean and a provided kernel (reference: gp_models/models.py:10-20).  The additive /
projected-additive per-component posterior models, the SVGP model and the RP->additive conversion of the reference
(:23-125) are not on the K.V hot path."""
from rpgp import gp as gpytorch
from rpgp.gp.models import ExactGP


class ExactGPModel(ExactGP):
    """Basic exact GP model with const mean and a provided kernel"""

    def __init__(self, train_x, train_y, likelihood, kernel):
        super(ExactGPModel, self).__init__(train_x, train_y, likelihood)
        self.mean_module = gpytorch.means.ConstantMean()
        self.covar_module = kernel

    def forward(self, x):
        mean_x = self.mean_module(x)
        covar_x = self.covar_module(x)
        return gpytorch.distributions.MultivariateNormal(mean_x, covar_x)
