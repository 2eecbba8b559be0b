"""Pivoted-Cholesky preconditioner and stochastic Lanczos quadrature: the pieces GPyTorch wraps around the K.V loop
(reached from the reference at training_routines.py:515,532 with the settings of
gp_experiment_runner.py:324-332).

The operator only has to serve `diag()` and single rows `rows([p])`.
"""
import torch


def pivoted_cholesky(op, max_iter, error_tol=1e-3):
    """Rank-<=max_iter pivoted Cholesky factor L (n x k) of the PSD operator `op` (K, not K + sigma^2 I).

    Greedy: pick the largest remaining diagonal entry, fetch that row of K, eliminate; stop after max_iter steps or
    when ||remaining diagonal||_1 / max(diag_0) < error_tol.  With a constant diagonal the first pivot is row 0.
    """
    n = op.shape[-1]
    diag = op.diag().clone()
    dtype, device = diag.dtype, diag.device
    orig_error = diag.max()
    errors = diag.norm(1) / orig_error
    L = torch.zeros(max_iter, n, dtype=dtype, device=device)
    perm = torch.arange(n, device=device)
    m = 0
    while (m == 0) or (m < max_iter and float(errors) > error_tol):
        permuted = diag[perm][m:]
        max_val, max_idx = permuted.max(0)
        max_idx = max_idx + m
        old = perm[m].clone()
        perm[m] = perm[max_idx]
        perm[max_idx] = old
        pi_m = perm[m]
        L[m, pi_m] = max_val.sqrt()
        row = op.rows(pi_m.view(1)).view(-1)
        if m + 1 < n:
            pi_i = perm[m + 1:]
            new = row[pi_i]
            if m > 0:
                new = new - (L[:m, pi_m].unsqueeze(-1) * L[:m][:, pi_i]).sum(0)
            new = new / L[m, pi_m]
            L[m, pi_i] = new
            cur = diag[pi_i] - new * new
            diag[pi_i] = cur
            errors = cur.norm(1) / orig_error
        m += 1
    return L[:m].t().contiguous()


class PivCholPreconditioner:
    """P = L L^T + sigma^2 I with L the pivoted-Cholesky factor (n x k).

    solve:  P^-1 v = (v - Q Q^T v) / sigma^2,  [L; sigma I_k] = Q R  (thin QR, Q restricted to its first n rows)
    logdet: 2 sum log|R_ii| + (n - k) log sigma^2
    sample: L e1 + sigma e2  ~ N(0, P)
    """

    def __init__(self, L, noise):
        n, k = L.shape
        self.L, self.noise = L, noise
        stacked = torch.cat([L, noise.sqrt() * torch.eye(k, dtype=L.dtype, device=L.device)], dim=0)
        Q, R = torch.linalg.qr(stacked)
        self.Q = Q[:n]
        self.logdet = R.diagonal().abs().log().sum() * 2 + (n - k) * noise.log()

    def solve(self, v):
        return (v - self.Q @ (self.Q.t() @ v)) / self.noise

    def sample(self, num, generator=None):
        n, k = self.L.shape
        e1 = torch.randn(k, num, dtype=self.L.dtype, device=self.L.device, generator=generator)
        e2 = torch.randn(n, num, dtype=self.L.dtype, device=self.L.device, generator=generator)
        return self.L @ e1 + self.noise.sqrt() * e2


def slq_logdet(T, n):
    """Stochastic Lanczos quadrature: log|A| ~= (n / p) sum_probes sum_k (e1^T q_k)^2 log(lambda_k) for the p
    tridiagonals T (p x k x k) that CG produced for p unit-norm probes; non-positive eigenvalues get weight 0."""
    evals, evecs = torch.linalg.eigh(T.double().cpu())
    weights = evecs[:, 0, :] ** 2
    mask = evals > 0
    logs = torch.where(mask, evals.clamp_min(1e-300).log(), torch.zeros_like(evals))
    per_probe = (weights * logs * mask).sum(-1)
    return float(n) * per_probe.mean()
