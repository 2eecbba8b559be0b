"""Kernel modules with GPyTorch's constructor / parameter names, lowered to the fused additive-RBF operator.

The reference assembles its covariance out of gpytorch.kernels pieces (RBFKernel, ScaleKernel, AdditiveKernel,
ProductKernel, AdditiveStructureKernel -- polynomial_projection_kernels.py:65-103, training_routines.py:148-174).
Here each of those classes is a thin parameter container: `forward(x1, x2)` does not evaluate anything, it returns an
`RPAdditiveLazyTensor` describing  sum_j c_j exp(-1/2 |z1_j - z2_j|^2)  (SURVEY.md §0), so that a whole kernel graph
collapses into ONE fused operator whose products run in the sm_100a kernels.  Parameter paths
(`kernel.kernels.0.base_kernel.kernels.1.raw_lengthscale`, `raw_outputscale`, ...) match GPyTorch's so state-dicts and
the reference's tests/utilities keep working (test.py:126-134, utils.py:12-29).
"""
import math

import torch

from ..lazy import LazyTensor, RPAdditiveLazyTensor
from .constraints import Positive
from .module import Module


def _same_points(x1, x2):
    return x2 is None or x2 is x1 or (x1.shape == x2.shape and torch.equal(x1, x2))


class LazyEvaluatedKernelTensor(LazyTensor):
    """What `kernel(x1, x2)` returns: nothing is computed until the solver asks (gpytorch.lazy.LazyEvaluatedKernelTensor,
    pinned by test.py:139-141)."""

    def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
        self.x1, self.x2, self.kernel = x1, x2, kernel
        self.last_dim_is_batch, self.params = last_dim_is_batch, params
        self._evaluated = None

    dtype = property(lambda self: self.x1.dtype)
    device = property(lambda self: self.x1.device)

    def _size(self):
        return torch.Size((self.x1.shape[-2], self.x2.shape[-2]))

    def evaluate_kernel(self):
        if self._evaluated is None:
            x2 = self.x1 if _same_points(self.x1, self.x2) else self.x2
            res = self.kernel.forward(self.x1, x2, diag=False, last_dim_is_batch=self.last_dim_is_batch, **self.params)
            if torch.is_tensor(res):
                from ..lazy import DenseLazyTensor
                res = DenseLazyTensor(res)
            self._evaluated = res
        return self._evaluated

    def representation(self):
        return self.evaluate_kernel().representation()

    def _matmul(self, rhs):
        return self.evaluate_kernel()._matmul(rhs)

    def matmul(self, rhs):
        return self.evaluate_kernel().matmul(rhs)

    def _quad_form_derivative(self, left_vecs, right_vecs):
        return self.evaluate_kernel()._quad_form_derivative(left_vecs, right_vecs)

    def diag(self):
        return self.evaluate_kernel().diag()

    def rows(self, index):
        return self.evaluate_kernel().rows(index)

    def evaluate(self):
        return self.evaluate_kernel().evaluate()

    def add_diag(self, diag):
        return self.evaluate_kernel().add_diag(diag)

    def _transpose_nonbatch(self):
        return self.evaluate_kernel()._transpose_nonbatch()


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        if len(batch_shape) != 0:
            raise NotImplementedError("batch kernels are outside the K.V hot path")
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor([active_dims] if isinstance(active_dims, int) else list(active_dims), dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps
        if self.has_lengthscale:
            num = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(1, num)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint or Positive())
            if lengthscale_prior is not None:
                self.register_prior("lengthscale_prior", lengthscale_prior, lambda m: m.lengthscale)

    # ---- lengthscale ---------------------------------------------------------------------------------------------------
    @property
    def lengthscale(self):
        if self.has_lengthscale:
            return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)
        return None

    @lengthscale.setter
    def lengthscale(self, value):
        if not self.has_lengthscale:
            raise RuntimeError("Kernel has no lengthscale.")
        self._set_constrained("raw_lengthscale", value)

    # ---- evaluation ----------------------------------------------------------------------------------------------------
    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_, x2_ = x1, x2
        if self.active_dims is not None:
            x1_ = x1_.index_select(-1, self.active_dims)
            if x2_ is not None:
                x2_ = x2_.index_select(-1, self.active_dims)
        if x1_.dim() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_ is not None:
            if x2_.dim() == 1:
                x2_ = x2_.unsqueeze(1)
            if x1_.shape[-1] != x2_.shape[-1]:
                raise RuntimeError("x1_ and x2_ must have the same number of dimensions!")
        if x2_ is None:
            x2_ = x1_
        if diag:
            res = self.forward(x1_, x2_, diag=True, last_dim_is_batch=last_dim_is_batch, **params)
            return res.diag() if isinstance(res, LazyTensor) else res
        return LazyEvaluatedKernelTensor(x1_, x2_, kernel=self, last_dim_is_batch=last_dim_is_batch, **params)

    def __add__(self, other):
        mine = list(self.kernels) if isinstance(self, AdditiveKernel) else [self]
        theirs = list(other.kernels) if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*(mine + theirs))

    def __mul__(self, other):
        mine = list(self.kernels) if isinstance(self, ProductKernel) else [self]
        theirs = list(other.kernels) if isinstance(other, ProductKernel) else [other]
        return ProductKernel(*(mine + theirs))


def _lower(kernel, x1, x2, last_dim_is_batch=False, **params):
    """Evaluate a sub-kernel the way gpytorch's composite kernels do (through __call__, so active_dims apply) and return
    its fused operator."""
    res = kernel(x1, x2, last_dim_is_batch=last_dim_is_batch, **params)
    res = res.evaluate_kernel() if isinstance(res, LazyTensor) else res
    if not isinstance(res, RPAdditiveLazyTensor):
        raise NotImplementedError(
            "%s does not lower to the additive-RBF operator; only RBF-based structures are on the K.V hot path"
            % kernel.__class__.__name__)
    return res


class RBFKernel(Kernel):
    """k(a, b) = exp(-1/2 |a/l - b/l|^2).  With last_dim_is_batch each input dimension becomes its own 1-D kernel
    (that is how AdditiveStructureKernel evaluates its base kernel)."""
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        same = _same_points(x1, x2)
        ls = self.lengthscale
        z1 = x1.div(ls)
        z2 = None if same else x2.div(ls)
        D = x1.shape[-1]
        one = torch.ones(1, dtype=x1.dtype, device=x1.device)
        if last_dim_is_batch:
            return RPAdditiveLazyTensor(z1, z2, one.expand(D), D, 1)
        return RPAdditiveLazyTensor(z1, z2, one, 1, D)


class MaternKernel(RBFKernel):
    """Matern nu = 1.5: (1 + sqrt3 d) exp(-sqrt3 d), d = |a/l - b/l| (gpytorch.kernels.MaternKernel as the reference selects it,
    training_routines.py:64-70: nu=1.5 only).  Same operator, base kernel 1 of the fused kernels."""
    _base = 1

    def __init__(self, nu=1.5, **kwargs):
        if nu != 1.5:
            raise NotImplementedError("only nu=1.5 (what the reference's _map_to_kernel builds) is on the fused K.V path")
        self.nu = nu
        super().__init__(**kwargs)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        op = super().forward(x1, x2, diag=diag, last_dim_is_batch=last_dim_is_batch, **params)
        op.base = self._base
        return op


class InverseMQKernel(MaternKernel):
    """(d^2 + 1)^-1/2 on lengthscale-divided inputs (gp_models/kernels/imq_kernel.py:8-22,25-58; both the dense and the KeOps
    class of the reference lower to this one operator)."""
    _base = 2

    def __init__(self, **kwargs):
        RBFKernel.__init__(self, **kwargs)


class CosineKernel(Kernel):
    """cos(pi |a - b| / period_length) (gpytorch.kernels.CosineKernel, which the reference's `_map_to_kernel` selects for
    kernel_type 'Cosine', training_routines.py:76-81; no lengthscale: `create_additive_rp_kernel` initialises period_length instead,
    :150-151).  Lowers to base kernel 3 of the fused kernels with pi / period_length folded into the coordinates."""
    has_lengthscale = False
    _base = 3

    def __init__(self, period_length_prior=None, period_length_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.register_parameter("raw_period_length", torch.nn.Parameter(torch.zeros(1, 1)))
        self.register_constraint("raw_period_length", period_length_constraint or Positive())
        if period_length_prior is not None:
            self.register_prior("period_length_prior", period_length_prior, lambda m: m.period_length)

    @property
    def period_length(self):
        return self.raw_period_length_constraint.transform(self.raw_period_length)

    @period_length.setter
    def period_length(self, value):
        self._set_constrained("raw_period_length", value)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        same = _same_points(x1, x2)
        scale = math.pi / self.period_length
        z1 = x1 * scale
        z2 = None if same else x2 * scale
        D = x1.shape[-1]
        one = torch.ones(1, dtype=x1.dtype, device=x1.device)
        op = RPAdditiveLazyTensor(z1, z2, one.expand(D), D, 1) if last_dim_is_batch else RPAdditiveLazyTensor(z1, z2, one, 1, D)
        op.base = self._base
        return op


class ScaleKernel(Kernel):
    """outputscale * base_kernel"""

    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        # GPyTorch's fix for the issue the reference works around by hand (polynomial_projection_kernels.py:87,94,99,146):
        # a ScaleKernel inherits the active dimensions of the kernel it wraps.
        if kwargs.get("active_dims") is None and getattr(base_kernel, "active_dims", None) is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(())))
        self.register_constraint("raw_outputscale", outputscale_constraint or Positive())
        if outputscale_prior is not None:
            self.register_prior("outputscale_prior", outputscale_prior, lambda m: m.outputscale)

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self._set_constrained("raw_outputscale", value)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        # like gpytorch, calls the base kernel's forward directly (its active_dims are NOT applied here, which is why
        # the reference passes active_dims to the ScaleKernel itself, polynomial_projection_kernels.py:87,94,99)
        res = self.base_kernel.forward(x1, x2, diag=False, last_dim_is_batch=last_dim_is_batch, **params)
        if isinstance(res, LazyTensor) and not isinstance(res, RPAdditiveLazyTensor):
            res = res.evaluate_kernel()
        if torch.is_tensor(res):
            return res * self.outputscale
        return res.scale(self.outputscale)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        return RPAdditiveLazyTensor.sum(_lower(k, x1, x2, last_dim_is_batch, **params) for k in self.kernels)


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        return RPAdditiveLazyTensor.product(_lower(k, x1, x2, last_dim_is_batch, **params) for k in self.kernels)


class AdditiveStructureKernel(Kernel):
    """sum over input dimensions of a 1-D base kernel (training_routines.py:169-171)"""

    def __init__(self, base_kernel, num_dims, active_dims=None):
        super().__init__(active_dims=active_dims)
        self.base_kernel = base_kernel
        self.num_dims = num_dims

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        if last_dim_is_batch:
            raise RuntimeError("AdditiveStructureKernel does not accept the last_dim_is_batch argument.")
        res = self.base_kernel.forward(x1, x2, diag=False, last_dim_is_batch=True, **params)
        if not isinstance(res, RPAdditiveLazyTensor):
            raise NotImplementedError("AdditiveStructureKernel needs an RBF-based base kernel on the K.V hot path")
        return res


class _KeOpsNamespace:
    """gpytorch.kernels.keops.* as the reference selects it with keops=True (training_routines.py:60-70): one backend here"""
    RBFKernel = RBFKernel
    MaternKernel = MaternKernel


keops = _KeOpsNamespace
