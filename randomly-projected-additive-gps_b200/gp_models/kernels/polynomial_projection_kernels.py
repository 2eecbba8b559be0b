"""Projection kernels of the rp_poly family: k(x, x') = sum_j s_j * prod_{m in group j} k1(p_m(x), p_m(x')).

API mirror of the reference's gp_models/kernels/polynomial_projection_kernels.py (class names, constructor arguments,
parameter paths such as `kernel.kernels.{j}.base_kernel.kernels.{m}.raw_lengthscale`, `kernel.kernels.{j}.raw_outputscale`,
`projection_module.weight/bias`; projection-cache semantics :119-137; `initialize` :139-156).  With an RBF base kernel a
product over a group is an ARD-RBF on the group's coordinates, so the whole object graph lowers to one fused additive-RBF
operator (SURVEY.md §0, row "rp_poly").  The SKI branch of the reference (:54-63, :77-82, :158-168) is an approximation
with a different algorithm and is not part of the exact K.V path: `ski=True` raises.
"""
import copy

import torch
from torch import nn

import rp
from gp_models.kernels.etc import _sample_from_range
from rpgp import gp as gpytorch
from rpgp import ops


class Identity(nn.Module):
    def forward(self, x):
        return x


def _no_ski(ski):
    if ski:
        raise NotImplementedError("SKI (GridInterpolationKernel) is an approximate path outside the exact K.V hot path")


class GeneralizedProjectionKernel(gpytorch.kernels.Kernel):
    """Sum over groups of products of one base kernel per projected coordinate.

    component_degrees=[2, 3] groups projected coordinates (0,1) and (2,3,4); every additive component has its own
    outputscale (trainable only when `weighted`), every coordinate its own lengthscale.
    """

    def __init__(self, component_degrees, d, base_kernel, projection_module, learn_proj=False, weighted=False, ski=False,
                 ski_options=None, X=None, lengthscale_prior=None, outputscale_prior=None, **kernel_kwargs):
        super(GeneralizedProjectionKernel, self).__init__()
        _no_ski(ski)
        self.learn_proj = learn_proj
        self.projection_module = projection_module
        self.ski = ski
        self.ski_options = ski_options
        self.base_kernel = base_kernel
        for param in self.projection_module.parameters():
            param.requires_grad = bool(self.learn_proj)

        n_groups = len(component_degrees)
        components = []
        coord = 0
        for degree in component_degrees:
            factors = []
            for _ in range(degree):
                factors.append(base_kernel(active_dims=coord, lengthscale_prior=copy.deepcopy(lengthscale_prior),
                                           **kernel_kwargs))
                coord += 1
            if degree == 1:
                inner = factors[0]
                scale_dims = inner.active_dims
            else:
                inner = gpytorch.kernels.ProductKernel(*factors)
                scale_dims = None
            prior = copy.deepcopy(outputscale_prior if weighted else lengthscale_prior)
            component = gpytorch.kernels.ScaleKernel(inner, outputscale_prior=prior, active_dims=scale_dims)
            component.initialize(outputscale=1 / n_groups)
            if not weighted:
                component.raw_outputscale.requires_grad = False
            components.append(component)

        self.kernel = gpytorch.kernels.AdditiveKernel(*components)
        self.kernel_kwargs = kernel_kwargs
        self.d = d
        self.component_degrees = component_degrees
        self.J = n_groups
        self.weighted = weighted
        self.last_x1 = None
        self.cached_projections = None
        self.cache_proj = not learn_proj  # may be switched off by hand

    def _project(self, x):
        """all projections of x at once: (n x d) -> (n x sum(component_degrees)).  A plain nn.Linear on float32 CUDA inputs goes through
        the library's tensor-core projection and its explicit vector-Jacobian product (rpgp.ops.project), the bias is added after."""
        pm = self.projection_module
        if isinstance(pm, torch.nn.Linear) and ops.can_project(x, pm.weight):
            z = ops.project(x, pm.weight)
            return z if pm.bias is None else z + pm.bias
        return self.projection_module(x)

    def forward(self, x1, x2, **params):
        if (not self.learn_proj) and self.cache_proj:
            last = self.last_x1
            if (last is not None and last.device == x1.device and last.dtype == x1.dtype and last.shape == x1.shape
                    and torch.equal(x1, last)):
                z1 = self.cached_projections
            else:
                z1 = self._project(x1)
                self.last_x1 = x1
                self.cached_projections = z1
        else:
            z1 = self._project(x1)
        same = x2 is x1 or torch.equal(x1, x2)
        z2 = z1 if same else self._project(x2)
        diag = params.pop("diag", False)
        params.pop("last_dim_is_batch", None)
        res = self.kernel(z1, z2, diag=diag, **params)
        return res if diag else res.evaluate_kernel()

    def initialize(self, mixin_range=None, lengthscale_range=None, **kwargs):
        """Sample the component weights (normalised to sum to 1) and one lengthscale per coordinate."""
        if mixin_range is None and lengthscale_range is None:
            return super(GeneralizedProjectionKernel, self).initialize(**kwargs)
        mixins = _sample_from_range(len(self.component_degrees), mixin_range)
        mixins = mixins / mixins.sum()
        for i, component in enumerate(self.kernel.kernels):
            component.outputscale = mixins[i]
            factors = [component.base_kernel] if self.component_degrees[i] == 1 else list(component.base_kernel.kernels)
            for factor in factors:
                factor.lengthscale = _sample_from_range(1, lengthscale_range)
        return self

    def to_additive_kernel(self):
        """A CustomAdditiveKernel over the projected coordinates that shares this kernel's component kernels."""
        groups, coord = [], 0
        for degree in self.component_degrees:
            groups.append(list(range(coord, coord + degree)))
            coord += degree
        res = CustomAdditiveKernel(groups, self.d, self.base_kernel, self.weighted, self.ski, self.ski_options,
                                   X=self.cached_projections, **self.kernel_kwargs)
        res.kernel = self.kernel
        return res

    @property
    def base_kernels(self):
        out = []
        for i, component in enumerate(self.kernel.kernels):
            if self.component_degrees[i] == 1:
                out.append(component.base_kernel)
            else:
                out.extend(component.base_kernel.kernels)
        return out

    @property
    def scale_kernels(self):
        return self.kernel.kernels


class GeneralizedPolynomialProjectionKernel(GeneralizedProjectionKernel):
    """J components of equal degree k"""

    def __init__(self, J, k, d, base_kernel, projection_module, learn_proj=False, weighted=False, ski=False,
                 ski_options=None, X=None, **kernel_kwargs):
        super(GeneralizedPolynomialProjectionKernel, self).__init__([k] * J, d, base_kernel, projection_module, learn_proj,
                                                                    weighted, ski, ski_options, X=X, **kernel_kwargs)
        self.J = J
        self.k = k


class PolynomialProjectionKernel(GeneralizedPolynomialProjectionKernel):
    """Linear projections given as lists: Ws[j] is (d x k), bs[j] is (k,)"""

    def __init__(self, J, k, d, base_kernel, Ws, bs, activation=None, learn_proj=False, weighted=False, ski=False,
                 ski_options=None, X=None, **kernel_kwargs):
        if activation is not None:
            raise ValueError("activation not supported through the normal projection interface. "
                             "Use the GeneralPolynomialProjectionKernel instead.")
        projection_module = torch.nn.Linear(d, J * k, bias=False)
        projection_module.weight = torch.nn.Parameter(torch.cat(Ws, dim=1).t())
        projection_module.bias = torch.nn.Parameter(torch.cat(bs, dim=0))
        super(PolynomialProjectionKernel, self).__init__(J, k, d, base_kernel, projection_module, learn_proj, weighted, ski,
                                                         ski_options, X=X, **kernel_kwargs)


class RPPolyKernel(PolynomialProjectionKernel):
    """Draws its own Gaussian projections (optionally diversified with rp.space_equally)"""

    def __init__(self, J, k, d, base_kernel, activation=None, learn_proj=False, weighted=False, space_proj=False,
                 ski=False, ski_options=None, X=None, **kernel_kwargs):
        projs = [rp.gen_rp(d, k) for _ in range(J)]
        bs = [torch.zeros(k) for _ in range(J)]
        if space_proj:
            newW, _ = rp.space_equally(torch.cat(projs, dim=1).t(), lr=0.1, niter=5000)
            newW.requires_grad = False
            projs = [newW[i:i + 1, :].t() for i in range(J)]
        super(RPPolyKernel, self).__init__(J, k, d, base_kernel, projs, bs, activation=activation, learn_proj=learn_proj,
                                           weighted=weighted, ski=ski, ski_options=ski_options, X=X, **kernel_kwargs)


class _GroupFeaturesModule(nn.Module):
    """'projection' that re-orders input features group by group"""

    def __init__(self, groups, d):
        super().__init__()
        order = [f for g in groups for f in g]
        self.register_buffer("order", torch.tensor(order))
        self.d = d

    def forward(self, x):
        return torch.index_select(x, -1, self.order)

    @property
    def weight(self):
        M = torch.zeros(self.d, self.d)
        for i, g in enumerate(self.order):
            M[i, g] = 1
        return M


class CustomAdditiveKernel(GeneralizedProjectionKernel):
    """Additive kernel over explicit feature groups"""

    def __init__(self, groups, d, base_kernel, weighted=False, ski=False, ski_options=None, X=None, **kernel_kwargs):
        kernel_kwargs.pop("learn_proj", None)  # the reference forwards it by accident
        super(CustomAdditiveKernel, self).__init__([len(g) for g in groups], d, base_kernel,
                                                   _GroupFeaturesModule(groups, d), weighted=weighted, ski=ski,
                                                   ski_options=ski_options, X=X, **kernel_kwargs)
        self.groups = groups


class StrictlyAdditiveKernel(CustomAdditiveKernel):
    """one 1-D kernel per input feature"""

    def __init__(self, d, base_kernel, weighted=False, ski=False, ski_options=None, X=None, **kernel_kwargs):
        super(StrictlyAdditiveKernel, self).__init__([[i] for i in range(d)], d, base_kernel, weighted=weighted, ski=ski,
                                                     ski_options=ski_options, X=X, **kernel_kwargs)
