"""Lanczos tridiagonalisation K ~= Q T Q^T and the low-rank root of K^-1 built from it -- what gpytorch's
`root_inv_decomposition` (method "lanczos") caches for `fast_pred_var` / LOVE (switched on by the reference
at gp_experiment_runner.py:235,327).  Every step is one K.V product of the fused operator."""
import torch


def lanczos_tridiag(matmul, init_vec, max_iter, tol=1e-5):
    """Q (n x k) with orthonormal columns and the tridiagonal T (k x k) with Q^T A Q = T; full re-orthogonalisation
    (k <= 100: n k^2 flops, negligible beside the k products).  Stops early when the residual norm falls below tol."""
    n = init_vec.shape[0]
    dtype, device = init_vec.dtype, init_vec.device
    q = init_vec.reshape(n) / init_vec.norm()
    Q = torch.zeros(n, max_iter, dtype=dtype, device=device)
    alpha = torch.zeros(max_iter, dtype=torch.float64)
    beta = torch.zeros(max_iter, dtype=torch.float64)
    k = 0
    for k in range(max_iter):
        Q[:, k] = q
        r = matmul(q.unsqueeze(-1)).squeeze(-1)
        a = torch.dot(q, r)
        alpha[k] = float(a)
        r = r - a * q
        if k > 0:
            r = r - float(beta[k - 1]) * Q[:, k - 1]
        # full re-orthogonalisation, twice is enough
        for _ in range(2):
            r = r - Q[:, :k + 1] @ (Q[:, :k + 1].t() @ r)
        b = float(r.norm())
        if k + 1 == max_iter or b < tol:
            k += 1
            break
        beta[k] = b
        q = r / b
    else:
        k = max_iter
    T = torch.diag(alpha[:k]) + torch.diag(beta[:k - 1], 1) + torch.diag(beta[:k - 1], -1)
    return Q[:, :k].contiguous(), T.to(dtype=torch.float64)


def lanczos_root_inv(matmul, init_vec, max_iter):
    """W (n x k) with A^-1 ~= W W^T on the Krylov space of init_vec:  A ~= Q T Q^T, T = L L^T, W = Q L^-T."""
    Q, T = lanczos_tridiag(matmul, init_vec, max_iter)
    jitter = 0.0
    eye = torch.eye(T.shape[0], dtype=T.dtype)
    for _ in range(6):
        try:
            L = torch.linalg.cholesky(T + jitter * eye)
            break
        except RuntimeError:
            jitter = 1e-8 * float(T.diagonal().mean()) if jitter == 0.0 else jitter * 10
    else:
        raise RuntimeError("lanczos_root_inv: tridiagonal matrix is not positive definite")
    Linv_t = torch.linalg.solve_triangular(L, eye, upper=False).t().to(dtype=Q.dtype, device=Q.device)
    return Q @ Linv_t
