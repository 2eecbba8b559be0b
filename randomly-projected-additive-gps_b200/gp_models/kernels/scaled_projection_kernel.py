"""ScaledProjectionKernel: the RPA-GP / DPA-GP kernel wrapper -- ARD scaling before (prescale) or after (postscale) a
linear projection, then a frozen additive base kernel.

Mirror of gp_models/kernels/scaled_projection_kernel.py:5-37 of the reference (same constructor, same parameter names:
`raw_lengthscale`, `projection_module.weight`, `base_kernel...`).  Instead of delegating to a GPyTorch kernel graph the
forward returns the fused operator over Z^ = proj(x / l) (or proj(x) / l), so every K.V of the solve runs in one
sm_100a kernel launch.
"""
import torch

from rpgp import gp as gpytorch
from rpgp import ops


class ScaledProjectionKernel(gpytorch.kernels.Kernel):
    def __init__(self, projection_module, base_kernel, prescale=False, ard_d=None, learn_proj=False, **kwargs):
        self.has_lengthscale = True
        super(ScaledProjectionKernel, self).__init__(ard_d=ard_d, **kwargs)
        self.projection_module = projection_module
        self.learn_proj = learn_proj
        if not self.learn_proj:
            for param in self.projection_module.parameters():
                param.requires_grad = False
        self.base_kernel = base_kernel
        for param in self.base_kernel.parameters():  # the additive base kernel stays frozen (test.py:597-599)
            param.requires_grad = False
        self.prescale = prescale

    def _scaled_projection(self, x):
        # float32 CUDA inputs through a bias-free nn.Linear take ONE library call (rpgp_project2_f32: tcgen05, 3xTF32) with an explicit
        # vector-Jacobian product for l and W (rpgp_project_bwd_f32) instead of a cuBLAS GEMM + element-wise kernels + autograd
        pm = self.projection_module
        if isinstance(pm, torch.nn.Linear) and pm.bias is None and ops.can_project(x, pm.weight):
            inv = self.lengthscale.reciprocal()
            return ops.project(x, pm.weight, inv, None) if self.prescale else ops.project(x, pm.weight, None, inv)
        if self.prescale:
            return self.projection_module(x.div(self.lengthscale))
        return self.projection_module(x).div(self.lengthscale)

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        same = x2 is x1 or torch.equal(x1, x2)
        z1 = self._scaled_projection(x1)
        z2 = z1 if same else self._scaled_projection(x2)
        res = self.base_kernel(z1, z2, diag=diag, last_dim_is_batch=last_dim_is_batch, **params)
        return res if diag else res.evaluate_kernel()
