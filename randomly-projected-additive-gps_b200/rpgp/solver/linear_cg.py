"""Preconditioned, multi-right-hand-side conjugate gradients with Lanczos tridiagonal extraction.

Restates the behaviour of GPyTorch's `linear_cg` (>=1.0; not vendored in the reference), the loop
that issues the K.V product on every iteration of an MLL evaluation (fitting/optimizing.py:65-74 ->
ExactMarginalLogLikelihood -> inv_quad_logdet -> linear_cg):

  * every right-hand-side column is normalised to unit 2-norm (zero columns keep norm 1) and un-normalised at the end;
  * x0 = 0; columns whose residual norm falls below `stop_updating_after` are frozen (alpha masked to 0);
  * divisions are guarded: denominators below `eps` give alpha = 0 / beta = 0;
  * stop when k >= min(10, max_iter-1) and mean_over_columns(||r||) < tolerance, but not before the Lanczos
    tridiagonals asked for have been collected (n_tridiag columns, at most max_tridiag_iter = 20 iterations);
  * T[k,k] = 1/alpha_k + beta_{k-1}/alpha_{k-1},  T[k,k-1] = T[k-1,k] = sqrt(beta_{k-1})/alpha_{k-1}; updating T stops once
    an off-diagonal falls below 1e-6.
"""
import warnings

import torch


class NumericalWarning(RuntimeWarning):
    pass


# running totals since import (bench.py reads them to report CG iterations per second)
STATS = {"solves": 0, "iterations": 0, "matmuls": 0}


def linear_cg(matmul_closure, rhs, n_tridiag=0, tolerance=1.0, eps=None, stop_updating_after=None, max_iter=1000,
              max_tridiag_iter=20, initial_guess=None, preconditioner=None, return_info=False):
    """Solve A X = rhs for the columns of rhs (n x t).  Returns X, or (X, T) with T (n_tridiag, k, k) when n_tridiag > 0.

    matmul_closure: callable V -> A V (the K^.V product), or a tensor.
    """
    if torch.is_tensor(matmul_closure):
        A = matmul_closure
        matmul_closure = A.matmul
    if eps is None:  # GPyTorch's guards (1e-10) cap the reachable residual near 1e-5; keep them in FP32, relax in FP64
        eps = 1e-10 if rhs.dtype == torch.float32 else 1e-24
    if stop_updating_after is None:
        stop_updating_after = 1e-10 if rhs.dtype == torch.float32 else 1e-14
    is_vector = rhs.dim() == 1
    if is_vector:
        rhs = rhs.unsqueeze(-1)
    n, t = rhs.shape[-2], rhs.shape[-1]
    if preconditioner is None:
        def preconditioner(x):
            return x
        precond = False
    else:
        precond = True
    zero_start = initial_guess is None      # x0 = 0: the first residual is the right-hand side itself -- no K.V product with zeros
    if zero_start:                          # (GPyTorch multiplies anyway; at n = 1M that is 2 s per solve for an exact zero)
        initial_guess = torch.zeros_like(rhs)

    n_iter = min(max_iter, n) if n > 0 else 0
    n_tridiag_iter = min(max_tridiag_iter, n)

    rhs_norm = rhs.norm(2, dim=-2, keepdim=True)
    rhs_is_zero = rhs_norm.lt(eps)
    rhs_norm = rhs_norm.masked_fill(rhs_is_zero, 1)
    rhs = rhs / rhs_norm

    residual = rhs.clone() if zero_start else rhs - matmul_closure(initial_guess)
    result = initial_guess.expand_as(residual).contiguous().clone()
    if not torch.equal(residual, residual):
        raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")

    residual_norm = residual.norm(2, dim=-2, keepdim=True)
    has_converged = residual_norm < stop_updating_after
    precond_residual = preconditioner(residual)
    curr_conjugate_vec = precond_residual
    residual_inner_prod = (precond_residual * residual).sum(-2, keepdim=True)

    if n_tridiag:
        t_mat = torch.zeros(n_tridiag_iter, n_tridiag_iter, n_tridiag, dtype=rhs.dtype, device=rhs.device)
        prev_alpha_recip = torch.empty(1, n_tridiag, dtype=rhs.dtype, device=rhs.device)
        prev_beta = torch.empty_like(prev_alpha_recip)
    update_tridiag = True
    last_tridiag_iter = 0
    tolerance_reached = False
    # convergence is tested on the host; on CUDA the mean residual norm travels through pinned memory and may be read `lag`
    # iterations late (settings.cg_convergence_lag) so that small, launch-bound solves do not synchronise every iteration
    lag = 0
    if rhs.is_cuda:
        from ..gp import settings as _s
        lag = _s.cg_convergence_lag.value()
        if lag < 0:
            lag = 0 if n * t >= (1 << 22) else 2
    pending = []          # (iteration, pinned scalar, event) of residual norms not yet read
    k = -1
    for k in range(n_iter):
        mvms = matmul_closure(curr_conjugate_vec)
        if precond:
            # preconditioned CG: alpha = <r, z> / <p, A p>
            pass
        denom = (curr_conjugate_vec * mvms).sum(-2, keepdim=True)
        is_small = denom < eps
        alpha = residual_inner_prod / denom.masked_fill(is_small, 1)
        alpha = alpha.masked_fill(is_small | has_converged, 0)

        result = result + alpha * curr_conjugate_vec
        residual = residual - alpha * mvms
        precond_residual = preconditioner(residual) if precond else residual

        new_inner = (residual * precond_residual).sum(-2, keepdim=True)
        is_small_b = residual_inner_prod < eps
        beta = new_inner / residual_inner_prod.masked_fill(is_small_b, 1)
        beta = beta.masked_fill(is_small_b, 0)
        residual_inner_prod = new_inner
        curr_conjugate_vec = precond_residual + beta * curr_conjugate_vec

        residual_norm = residual.norm(2, dim=-2, keepdim=True).masked_fill(rhs_is_zero, 0)
        has_converged = residual_norm < stop_updating_after

        if n_tridiag and k < n_tridiag_iter and update_tridiag:
            alpha_t = alpha[..., :n_tridiag]
            beta_t = beta[..., :n_tridiag]
            alpha_recip = 1.0 / alpha_t.masked_fill(alpha_t == 0, 1)  # frozen columns contribute 1 (as GPyTorch does)
            if k == 0:
                t_mat[0, 0] = alpha_recip[0]
            else:
                t_mat[k, k] = (alpha_recip + prev_beta * prev_alpha_recip)[0]
                off = (prev_beta.sqrt() * prev_alpha_recip)[0]
                t_mat[k, k - 1] = off
                t_mat[k - 1, k] = off
                if float(off.max()) < 1e-6:
                    update_tridiag = False
            last_tridiag_iter = k
            prev_alpha_recip = alpha_recip.clone()
            prev_beta = beta_t.clone()

        if k >= min(10, max_iter - 1) and not (n_tridiag and k < min(n_tridiag_iter, max_iter - 1)):
            if lag == 0:
                if float(residual_norm.mean()) < tolerance:
                    tolerance_reached = True
                    break
            else:
                host = torch.empty((), dtype=residual_norm.dtype, pin_memory=True)
                host.copy_(residual_norm.mean(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((k, host, ev))
                if len(pending) > lag:
                    _, h, e = pending.pop(0)
                    e.synchronize()
                    if float(h) < tolerance:
                        tolerance_reached = True
                        break
    if lag and not tolerance_reached and pending:      # max_iter reached with unread norms: the last one decides
        _, h, e = pending[-1]
        e.synchronize()
        tolerance_reached = float(h) < tolerance

    STATS["solves"] += 1
    STATS["iterations"] += k + 1
    STATS["matmuls"] += k + 1 + (0 if zero_start else 1)
    result = result * rhs_norm
    if not tolerance_reached and n_iter > 0:
        warnings.warn(
            "CG terminated in {} iterations with average residual norm {} which is larger than the tolerance of {} "
            "specified by settings.cg_tolerance. If performance is affected, consider raising the maximum number of CG "
            "iterations by running code in a settings.max_cg_iterations(value) context.".format(
                k + 1, float(residual_norm.mean()), tolerance), NumericalWarning)
    if is_vector:
        result = result.squeeze(-1)
    info = {"iterations": k + 1, "residual_norm": float(residual_norm.mean()) if n_iter > 0 else 0.0,
            "converged": tolerance_reached}
    if n_tridiag:
        T = t_mat[:last_tridiag_iter + 1, :last_tridiag_iter + 1].permute(2, 0, 1).contiguous()
        return (result, T, info) if return_info else (result, T)
    return (result, info) if return_info else result
