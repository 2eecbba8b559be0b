"""Exact GP models of the reference (gp_models/models.py): the basic model with a constant mean and a provided kernel (:10-20), the
additive model with per-component posteriors `additive_pred` (:23-72), the projected-additive model and the RP -> additive
conversion (:75-86,110-125).  The SVGP model (:89-107) is a variational family outside the exact K.V path.

Per-component posteriors reuse the fused operator: component j of the additive kernel is the single-group operator
c_j k(z^(j), z'^(j)), so the J component means cost J rectangular products of one group each -- the exponentials of ONE full K.V --
after a single solve K^-1 (y - mu); component covariances stay lazy (rpgp/lazy.py PredictiveCovarLazyTensor).
"""
import torch

from rpgp import gp as gpytorch
from rpgp import lazy
from rpgp.gp.models import ExactGP

from .kernels import CustomAdditiveKernel, GeneralizedProjectionKernel


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


def _unwrap(kernel, cls, what):
    """(outputscale or None, inner kernel) of `kernel` or `ScaleKernel(kernel)`; ValueError when it is not a `cls`"""
    inner = kernel.base_kernel if isinstance(kernel, gpytorch.kernels.ScaleKernel) else kernel
    if not isinstance(inner, cls):
        raise ValueError("Not %s kernel." % what)
    return (kernel.outputscale if inner is not kernel else None), inner


class AdditiveExactGPModel(ExactGPModel):
    def __init__(self, train_x, train_y, likelihood, kernel):
        _unwrap(kernel, CustomAdditiveKernel, "an additive")
        super(AdditiveExactGPModel, self).__init__(train_x, train_y, likelihood, kernel)

    def additive_pred(self, x, group=None):
        """Posterior of every additive component (or of component `group`) at x: MultivariateNormal(K_j* K^-1 (y - mu),
        K_j** - K_j* K^-1 K_j*^T) with K_j the component's kernel times the outer outputscale.  The component means add up to
        the model's predictive mean minus its prior mean (test.py:403-405).  Two deliberate differences from the reference
        (:33-64): the prior mean is subtracted from the targets (the reference solves against y itself, which is the same thing
        only for a zero mean), and a component acts on the re-grouped features its active_dims refer to (the reference hands it the
        raw inputs, which coincides for groups given in feature order)."""
        scale, add_kernel = _unwrap(self.covar_module, CustomAdditiveKernel, "an additive")
        train_x = self.train_inputs[0]
        with torch.no_grad():
            prior = self.forward(train_x)
            train_covar = self.likelihood(prior).lazy_covariance_matrix          # K + sigma^2 I, all components
            K_inv_y = train_covar.inv_matmul((self.train_targets - prior.mean).unsqueeze(-1))
            z_test, z_train = add_kernel._project(x), add_kernel._project(train_x)

            def get_pred(component):
                cross = component(z_test, z_train).evaluate_kernel()
                test_test = component(z_test, z_test).evaluate_kernel()
                if scale is not None:
                    cross, test_test = cross.scale(scale), test_test.scale(scale)
                mean = cross._matmul(K_inv_y).squeeze(-1)
                return gpytorch.distributions.MultivariateNormal(mean, lazy.PredictiveCovarLazyTensor(test_test, cross, train_covar))

            components = add_kernel.kernel.kernels
            if group is None:
                return [get_pred(k) for k in components]
            return get_pred(components[group])

    def get_groups(self):
        return _unwrap(self.covar_module, CustomAdditiveKernel, "an additive")[1].groups


class ProjectedAdditiveExactGPModel(ExactGPModel):
    def __init__(self, train_x, train_y, likelihood, kernel):
        _unwrap(kernel, GeneralizedProjectionKernel, "a projected additive")
        super(ProjectedAdditiveExactGPModel, self).__init__(train_x, train_y, likelihood, kernel)

    def get_corresponding_additive_model(self, return_proj=True):
        return convert_rp_model_to_additive_model(self, return_proj=return_proj)


def convert_rp_model_to_additive_model(model, return_proj=True):
    """The additive model over the PROJECTED training inputs that shares the RP model's component kernels, likelihood and mean
    (reference :110-125): its predictions at projection(x) equal the RP model's at x (test.py:359-380)."""
    scale, rp_kernel = _unwrap(model.covar_module, GeneralizedProjectionKernel, "a projected additive")
    add_kernel = rp_kernel.to_additive_kernel()
    proj = rp_kernel.projection_module
    if scale is not None:
        add_kernel = gpytorch.kernels.ScaleKernel(add_kernel)
        add_kernel.initialize(outputscale=scale.detach())
    with torch.no_grad():
        Z = rp_kernel.projection_module(model.train_inputs[0])
    res = AdditiveExactGPModel(Z, model.train_targets, model.likelihood, add_kernel)
    res.mean_module = model.mean_module
    res = res.to(Z.device, Z.dtype)
    return (res, proj) if return_proj else res
