"""inv_quad_logdet: (y^T K^-1 y, log|K|) and their gradients, the body of ExactMarginalLogLikelihood.

Restates GPyTorch's `InvQuadLogDet` function (reached from fitting/optimizing.py:65-74):

  forward   rhs = [z_1 .. z_p | y - mu], z ~ N(0, P) (or N(0, I) without preconditioner), unit-normalised;
            one multi-RHS preconditioned CG solve (the K.V hot loop); logdet from the p Lanczos tridiagonals
            (stochastic Lanczos quadrature) + log|P|; inv_quad = (y - mu)^T x_y.
  backward  L = [ (1/p) |z| K^-1 z^ g_logdet | -g_iq x_y ],  R = [ P^-1 (|z| z^) | x_y ]
            -> operator._quad_form_derivative(L, R)    (one fused gradient kernel pass)
            grad wrt y - mu = 2 g_iq x_y.

Small problems (n <= settings.max_cholesky_size, or fast_computations off / --use_chol) take the dense Cholesky route
with ordinary autograd through the differentiable dense evaluation.
"""
import torch

from .. import dist as rdist
from .preconditioner import slq_logdet


def _settings():
    from ..gp import settings
    return settings


class _InvQuadLogDetCG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op_template, compute_logdet, rhs, *rep):
        s = _settings()
        op = op_template._rebuild(*[r.detach() for r in rep])
        n = op.shape[-1]
        if rdist.world_size() > 1:
            rdist.assert_replicated("the MLL solve's right-hand side / operator representation", rhs, *[r for r in rep if torch.is_tensor(r)])
        dtype, device = rhs.dtype, rhs.device
        precond = op._preconditioner()
        num_probes = s.num_trace_samples.value() if compute_logdet else 0
        probes = norms = None
        if num_probes > 0:
            fixed = s.deterministic_probes.probe_vectors
            if fixed is not None and fixed.shape == (n, num_probes):
                probes = fixed.to(dtype=dtype, device=device).clone()
            elif precond is None:
                probes = torch.randn(n, num_probes, dtype=dtype, device=device)
            else:
                probes = precond.sample(num_probes).to(dtype)
            if rdist.world_size() > 1:      # one set of probes for the whole job (local RNG streams differ)
                probes = rdist.broadcast_(probes.contiguous())
            norms = probes.norm(2, dim=-2, keepdim=True)
            probes = probes / norms
            full_rhs = torch.cat([probes, rhs], dim=-1)
        else:
            full_rhs = rhs
        if num_probes > 0:
            solves, T = op._solve(full_rhs, precond, num_tridiag=num_probes)
        else:
            solves, T = op._solve(full_rhs, precond), None
        iq_solves = solves[:, num_probes:]
        inv_quad = (iq_solves * rhs).sum()
        logdet = torch.zeros((), dtype=dtype, device=device)
        if num_probes > 0:
            if not s.skip_logdet_forward.on():
                logdet = slq_logdet(T, n).to(dtype=dtype, device=device)
                if precond is not None:
                    logdet = logdet + precond.logdet.to(dtype)
        ctx.op, ctx.precond, ctx.num_probes = op, precond, num_probes
        ctx.rep_requires = [r.requires_grad for r in rep]
        ctx.save_for_backward(solves, probes if probes is not None else rhs.new_zeros(0), norms if norms is not None else rhs.new_zeros(0))
        return inv_quad, logdet

    @staticmethod
    def backward(ctx, g_iq, g_ld):
        solves, probes, norms = ctx.saved_tensors
        p = ctx.num_probes
        iq_solves = solves[:, p:]
        lefts, rights = [], []
        if p > 0:
            coef = 1.0 / p
            lefts.append(solves[:, :p] * (coef * norms * g_ld))
            pv = probes * norms
            if ctx.precond is not None:
                pv = ctx.precond.solve(pv)
            rights.append(pv)
        lefts.append(-iq_solves * g_iq)
        rights.append(iq_solves)
        L = torch.cat(lefts, dim=-1)
        R = torch.cat(rights, dim=-1)
        grads = ctx.op._quad_form_derivative(L, R)
        grads = tuple(g if need else None for g, need in zip(grads, ctx.rep_requires))
        rhs_grad = 2.0 * iq_solves * g_iq
        return (None, None, rhs_grad) + grads


def psd_safe_cholesky(A, max_tries=3):
    """Cholesky factor of A; when the factorisation fails, retried with a diagonal jitter of 1e-6 (1e-8 in FP64) times 1, 10, 100.
    GPyTorch's `psd_safe_cholesky` (gpytorch/utils/cholesky.py, the routine behind every dense root of the reference's models):
    a predictive covariance assembled from CG solves at eval_cg_tolerance can miss positive definiteness by rounding."""
    try:
        return torch.linalg.cholesky(A)
    except RuntimeError as err:
        if torch.isnan(A).any():
            raise
        first = err
    jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    eye = torch.eye(A.shape[-1], dtype=A.dtype, device=A.device)
    for i in range(max_tries):
        try:
            return torch.linalg.cholesky(A + (jitter * 10.0 ** i) * eye)
        except RuntimeError:
            continue
    raise first


def inv_quad_logdet(op, inv_quad_rhs=None, logdet=False):
    """op: AddedDiagLazyTensor (K + sigma^2 I).  Returns (inv_quad, logdet) as 0-dim tensors (None when not asked)."""
    n = op.shape[-1]
    if inv_quad_rhs is None:
        rhs = torch.zeros(n, 1, dtype=op.dtype, device=op.device)
    else:
        rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
    if op._use_cholesky():
        Kd = op.evaluate()
        Lc = psd_safe_cholesky(Kd)
        iq = None
        if inv_quad_rhs is not None:
            sol = torch.linalg.solve_triangular(Lc, rhs, upper=False)
            iq = (sol * sol).sum()
        ld = 2.0 * Lc.diagonal().log().sum() if logdet else None
        return iq, ld
    iq, ld = _InvQuadLogDetCG.apply(op, bool(logdet), rhs, *op.representation())
    return (iq if inv_quad_rhs is not None else None), (ld if logdet else None)
