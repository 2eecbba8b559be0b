"""Memory-efficient GAM kernel: K = sum_i exp(-1/2 ((x1_i - x2_i) / l_i)^2), the DPA-GP base kernel.

Reference: gp_models/kernels/memory_efficient_gam_kernel.py.  There `GAMFunction` (:5-59) is a torch.autograd.Function
that accumulates the d one-dimensional RBF kernels into a dense n x m matrix with a hand-written backward.  Here the
same Function signature -- forward(ctx, x1, x2, lengthscale) -> K, backward -> (x1_grad, x2_grad, lengthscale_grad) --
is served by the CUDA kernels (dense rows forward; the row-gradient kernel with L = grad_output, R = I backward), and the
Kernel class returns the fused lazy operator instead of the dense matrix so that products never materialise K.
"""
import torch

from rpgp import gp as gpytorch
from rpgp import ops
from rpgp.lazy import RPAdditiveLazyTensor


class GAMFunction(torch.autograd.Function):
    """Dense sum of 1-D RBF kernels (n x m).  Does not support batch mode (neither does the reference)."""

    @staticmethod
    def forward(ctx, x1, x2, lengthscale):
        n, d = x1.shape
        m, d2 = x2.shape
        if d2 != d:
            raise ValueError("Dimension mismatch")
        ctx.save_for_backward(x1, x2, lengthscale)
        ls = lengthscale.reshape(1, -1)
        z1 = x1.detach().div(ls).contiguous()
        z2 = x2.detach().div(ls).contiguous()
        ones = torch.ones(d, dtype=x1.dtype, device=x1.device)
        return ops.kernel_rows_raw(z1, z2, ones, d, 1)

    @staticmethod
    def backward(ctx, grad_output):
        x1, x2, lengthscale = ctx.saved_tensors
        n, d = x1.shape
        m = x2.shape[0]
        ls = lengthscale.reshape(1, -1).detach()
        z1 = x1.detach().div(ls).contiguous()
        z2 = x2.detach().div(ls).contiguous()
        ones = torch.ones(d, dtype=x1.dtype, device=x1.device)
        dz1 = torch.zeros_like(z1)
        dz2 = torch.zeros_like(z2)
        eye = torch.eye(m, dtype=x1.dtype, device=x1.device)
        G = grad_output.contiguous()
        for c0 in range(0, m, 16):  # S = grad_output = sum over column blocks of G[:, blk] I[:, blk]^T
            c1 = min(m, c0 + 16)
            a, b, _ = ops.quad_form_grads(z1, z2, ones, d, 1, G[:, c0:c1].contiguous(), eye[:, c0:c1].contiguous(), False)
            dz1 += a
            dz2 += b
        x1_grad = dz1 / ls if ctx.needs_input_grad[0] else None
        x2_grad = dz2 / ls if ctx.needs_input_grad[1] else None
        # z = x / l  =>  d/dl = -sum_rows dz * z / l   (per dimension; summed when there is a single lengthscale)
        per_dim = -((dz1 * z1).sum(0) + (dz2 * z2).sum(0)) / ls.reshape(-1)
        if lengthscale.numel() == 1:
            ls_grad = per_dim.sum().reshape(lengthscale.shape)
        else:
            ls_grad = per_dim.reshape(lengthscale.shape)
        return x1_grad, x2_grad, ls_grad


class MemoryEfficientGamKernel(gpytorch.kernels.Kernel):
    def __init__(self, **kwargs):
        self.has_lengthscale = True
        super(MemoryEfficientGamKernel, self).__init__(has_lengthscale=True, **kwargs)
        self.covar_dist = GAMFunction  # `.apply` is static; instantiating autograd Functions is deprecated

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
        same = x2 is x1 or torch.equal(x1, x2)
        ls = self.lengthscale
        z1 = x1.div(ls)
        z2 = None if same else x2.div(ls)
        d = x1.shape[-1]
        op = RPAdditiveLazyTensor(z1, z2, torch.ones(d, dtype=x1.dtype, device=x1.device), d, 1)
        return op.diag() if diag else op
