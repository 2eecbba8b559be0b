"""CPU oracle for the K.V hot path of Randomly-Projected-Additive-GPs.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy (FP64 by default) restatement of the reference's algorithm for the path
named in BASELINE.json `north_star`.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it -- the product path
(`randomly-projected-additive-gps_b200/rpgp`) never does and fails loudly without its CUDA library.

Parity pinning:
  * pinned by the reference's own known-answer tests G1-G5 (test.py:533-573, 625-635, 640-680), and by
    fixtures generated in the build container by importing the reference's `rp.py` and the class body of
    `GAMFunction` (tests/golden/make_golden.py -> tests/golden/*.npz);
  * "parity unpinned" by any reference test for: CG solutions, MLL values, MLL gradients, predictions.
    For those the oracle below is the dense FP64 Cholesky restatement of GPyTorch's exact-GP algebra
    (SURVEY.md Appendix A); GPyTorch itself (>=1.0, un-pinned, README.md:7) and PyKeOps (>=1.2, README.md:8)
    are third-party dependencies that are absent from /root/reference and not installable here.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Arithmetic uses direct differences (no |a|^2+|b|^2-2ab trick) so FP64 results are good to ~1e-15.
"""
import math

import numpy as np

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


# ----------------------------------------------------------------------------------------------
# parametrisation helpers (GPyTorch Positive() constraint = softplus)
# ----------------------------------------------------------------------------------------------
def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 30, x, np.log1p(np.exp(np.minimum(x, 30))))


def inv_softplus(y):
    y = np.asarray(y, dtype=np.float64)
    return np.where(y > 30, y, np.log(np.expm1(np.minimum(y, 30))))


# ----------------------------------------------------------------------------------------------
# projection: scaled_projection_kernel.py:21-37 (ScaledProjectionKernel.forward)
# ----------------------------------------------------------------------------------------------
def scaled_projection(X, W, ell, prescale, dtype=np.float64):
    """Z = (X/ell) W^T when prescale (ell has d entries), else (X W^T)/ell (ell has J*K entries).

    W is the torch.nn.Linear weight, shape (J*K, d)  (training_routines.py:144-145).
    """
    X = np.asarray(X, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    ell = np.asarray(ell, dtype=dtype).reshape(1, -1)
    if prescale:
        return (X / ell) @ W.T
    return (X @ W.T) / ell


# ----------------------------------------------------------------------------------------------
# the canonical operator:  K[i,i'] = sum_j c_j exp(-1/2 sum_m (Z1[i,jK+m]-Z2[i',jK+m])^2)
# restates AdditiveKernel(ScaleKernel(RBF(active_dims=group j)))  -- polynomial_projection_kernels.py:65-103,
# training_routines.py:148-174 -- with RBF = exp(-1/2 |a-b|^2) (KeOps form, imq_kernel.py:44-47 analogue).
# ----------------------------------------------------------------------------------------------
# base kernels of a group as functions of its squared distance (training_routines.py:57-83 `_map_to_kernel`):
#   0 RBF  exp(-sq/2)            gpytorch RBFKernel / keops RBFKernel
#   1 Matern nu=1.5  (1 + sqrt(3 sq)) exp(-sqrt(3 sq))   gpytorch MaternKernel(nu=1.5) [GPyTorch, recalled]
#   2 inverse multiquadric  (sq + 1)^-1/2   gp_models/kernels/imq_kernel.py:8-9 (postprocess_inverse_mq), :47 (KeOps form)
#   3 cosine  cos(sqrt(sq))   gpytorch CosineKernel, cos(pi |a - b| / period_length) [GPyTorch, recalled; training_routines.py:76-81,
#     :150-151]: the kernel class folds pi / period_length into the coordinates
def base_f(base, sq):
    if base == 1:
        q = np.sqrt(3.0 * sq)
        return (1.0 + q) * np.exp(-q)
    if base == 2:
        return 1.0 / np.sqrt(sq + 1.0)
    if base == 3:
        return np.cos(np.sqrt(sq))
    return np.exp(-0.5 * sq)


def base_df(base, sq):
    """d f / d sq"""
    if base == 1:
        return -1.5 * np.exp(-np.sqrt(3.0 * sq))
    if base == 2:
        return -0.5 * (sq + 1.0) ** -1.5
    if base == 3:        # -sin(d) / (2 d), -> -1/2 at the origin
        d = np.sqrt(sq)
        with np.errstate(divide="ignore", invalid="ignore"):
            q = -0.5 * np.sin(d) / d
        return np.where(d < 1e-6, -0.5 + sq / 12.0, q)
    return -0.5 * np.exp(-0.5 * sq)


def additive_rbf_dense(Z1, Z2, c, J, K, dtype=np.float64, base=0):
    Z1 = np.asarray(Z1, dtype=dtype)
    Z2 = np.asarray(Z2, dtype=dtype)
    c = np.broadcast_to(np.asarray(c, dtype=dtype), (J,))
    m, n = Z1.shape[0], Z2.shape[0]
    out = np.zeros((m, n), dtype=dtype)
    for j in range(J):
        sq = np.zeros((m, n), dtype=dtype)
        for q in range(j * K, (j + 1) * K):
            diff = Z1[:, q][:, None] - Z2[:, q][None, :]
            sq += diff * diff
        out += c[j] * base_f(base, sq)
    return out


def kmv(Z1, Z2, c, J, K, V, diag_add=0.0, row_chunk=1024, dtype=np.float64, base=0):
    """out = K(Z1,Z2) @ V (+ diag_add * V when square); row-chunked like gpytorch checkpoint_kernel
    (gp_experiment_runner.py:250,330) so that K is never held whole."""
    Z1 = np.asarray(Z1, dtype=dtype)
    V = np.asarray(V, dtype=dtype)
    out = np.empty((Z1.shape[0], V.shape[1]), dtype=dtype)
    for r0 in range(0, Z1.shape[0], row_chunk):
        r1 = min(r0 + row_chunk, Z1.shape[0])
        out[r0:r1] = additive_rbf_dense(Z1[r0:r1], Z2, c, J, K, dtype=dtype, base=base) @ V
    if diag_add != 0.0:
        out += diag_add * V
    return out


def kernel_rows(Z1, Z2, c, J, K, rows, dtype=np.float64):
    """K[rows, :] -- the pivoted-Cholesky row fetch."""
    return additive_rbf_dense(np.asarray(Z1, dtype=dtype)[np.asarray(rows)], Z2, c, J, K, dtype=dtype)


def kernel_diag(c, J, n, dtype=np.float64):
    """diag K(Z,Z) = sum_j c_j (constant)."""
    return np.full((n,), np.broadcast_to(np.asarray(c, dtype=dtype), (J,)).sum(), dtype=dtype)


# ----------------------------------------------------------------------------------------------
# quadratic-form derivative:  G = sum_col L[:,col]^T K R[:,col]
# explicit per-pair formulas, the analogue of GAMFunction.backward memory_efficient_gam_kernel.py:33-59
# ----------------------------------------------------------------------------------------------
def quad_form_grads(Z1, Z2, c, J, K, L, R, dtype=np.float64, base=0):
    """Return (dG/dZ1, dG/dZ2, dG/dc) with Z1, Z2 treated as independent tensors."""
    Z1 = np.asarray(Z1, dtype=dtype)
    Z2 = np.asarray(Z2, dtype=dtype)
    L = np.asarray(L, dtype=dtype)
    R = np.asarray(R, dtype=dtype)
    c = np.broadcast_to(np.asarray(c, dtype=dtype), (J,))
    S = L @ R.T  # (m, n) weights dG/dK
    dZ1 = np.zeros_like(Z1)
    dZ2 = np.zeros_like(Z2)
    dc = np.zeros((J,), dtype=dtype)
    for j in range(J):
        sq = np.zeros_like(S)
        diffs = []
        for q in range(j * K, (j + 1) * K):
            diff = Z1[:, q][:, None] - Z2[:, q][None, :]
            diffs.append(diff)
            sq += diff * diff
        dc[j] = (S * base_f(base, sq)).sum()
        w = -2.0 * S * base_df(base, sq) * c[j]                  # RBF: S k c
        for mm, q in enumerate(range(j * K, (j + 1) * K)):
            wd = w * diffs[mm]
            dZ1[:, q] = -wd.sum(axis=1)
            dZ2[:, q] = wd.sum(axis=0)
    return dZ1, dZ2, dc


# ----------------------------------------------------------------------------------------------
# GAMFunction restatement: memory_efficient_gam_kernel.py:12-30 (forward), :33-59 (backward)
# ----------------------------------------------------------------------------------------------
def gam_forward(x1, x2, lengthscale, dtype=np.float64):
    x1 = np.asarray(x1, dtype=dtype)
    x2 = np.asarray(x2, dtype=dtype)
    ls = np.asarray(lengthscale, dtype=dtype).reshape(-1)
    n, d = x1.shape
    if x2.shape[1] != d:
        raise ValueError("Dimension mismatch")  # memory_efficient_gam_kernel.py:15-16
    a = x1 / ls
    b = x2 / ls
    kernel = np.zeros((n, x2.shape[0]), dtype=dtype)
    for i in range(d):
        diff = a[:, i][:, None] - b[:, i][None, :]
        kernel += np.exp(-0.5 * diff * diff)
    return kernel


def gam_backward(x1, x2, lengthscale, grad_output, dtype=np.float64):
    """(x1_grad, x2_grad, lengthscale_grad) for upstream grad_output (n x m).  :33-59."""
    x1 = np.asarray(x1, dtype=dtype)
    x2 = np.asarray(x2, dtype=dtype)
    g = np.asarray(grad_output, dtype=dtype)
    ls = np.asarray(lengthscale, dtype=dtype).reshape(-1)
    num_l = ls.size
    n, d = x1.shape
    lsd = np.broadcast_to(ls, (d,)) if num_l == 1 else ls
    a = x1 / lsd
    b = x2 / lsd
    x1g = np.zeros_like(x1)
    x2g = np.zeros_like(x2)
    lsg = np.zeros((num_l,), dtype=dtype)
    for i in range(d):
        diff = b[:, i][None, :] - a[:, i][:, None]          # x2_ - x1_  (:45)
        sq = diff * diff
        dk = np.exp(-0.5 * sq) * g                           # Delta_K (:47)
        idx = i if num_l > 1 else 0
        lsg[idx] += (sq * dk).sum() / lsd[i]                 # (:49)
        dkd = diff * dk                                      # (:52)
        x1g[:, i] = dkd.sum(axis=1) / lsd[i]                 # (:54)
        x2g[:, i] = -dkd.sum(axis=0) / lsd[i]                # (:56)
    return x1g, x2g, lsg


# ----------------------------------------------------------------------------------------------
# exact (dense Cholesky) marginal log likelihood + gradients: SURVEY Appendix A
#   ExactMLL = [ -1/2 ( r^T Khat^-1 r + log|Khat| + n log 2pi ) + sum_priors ] / n,  r = y - mu
#   call sites: fitting/optimizing.py:65-74, training_routines.py:515,532
# ----------------------------------------------------------------------------------------------
def smoothed_box_log_prob(x, a=1e-4, b=10.0, sigma=0.01):
    """gpytorch.priors.SmoothedBoxPrior(a, b, sigma) log-density at x (training_routines.py:345-350)."""
    cc = 0.5 * (a + b)
    rr = 0.5 * (b - a)
    m_ = 1.0 + (b - a) / (math.sqrt(2.0 * math.pi) * sigma)
    xt = max(abs(x - cc) - rr, 0.0)
    normal_lp = -0.5 * (xt / sigma) ** 2 - math.log(sigma) - 0.5 * math.log(2.0 * math.pi)
    return normal_lp - math.log(m_)


def exact_mll_dense(Khat, y, mean_const=0.0, log_prior=0.0):
    """Exact MLL value (divided by n) from a dense Khat = K + noise I."""
    Khat = np.asarray(Khat, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = y.size
    r = y - mean_const
    Lc = np.linalg.cholesky(Khat)
    alpha = np.linalg.solve(Lc.T, np.linalg.solve(Lc, r))
    logdet = 2.0 * np.log(np.diag(Lc)).sum()
    inv_quad = r @ alpha
    mll = -0.5 * (inv_quad + logdet + n * math.log(2.0 * math.pi)) + log_prior
    return mll / n, alpha, logdet, inv_quad


def exact_mll_grads_dense(Z, c, J, K, noise, y, mean_const=0.0):
    """d(n * MLL)/d{Z, c, noise, mean} for Khat = K(Z,Z) + noise I (priors excluded), dense FP64.

    Uses dMLL/dKhat = 1/2 (alpha alpha^T - Khat^-1); the symmetric operand is counted in both roles
    (scaled_projection_kernel.py:29-30 sets x2 = x1).
    """
    Z = np.asarray(Z, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = y.size
    Kd = additive_rbf_dense(Z, Z, c, J, K)
    Khat = Kd + noise * np.eye(n)
    Kinv = np.linalg.inv(Khat)
    r = y - mean_const
    alpha = Kinv @ r
    G = 0.5 * (np.outer(alpha, alpha) - Kinv)  # d(n*MLL)/dKhat
    cc = np.broadcast_to(np.asarray(c, dtype=np.float64), (J,))
    dZ = np.zeros_like(Z)
    dc = np.zeros((J,))
    for j in range(J):
        sq = np.zeros((n, n))
        diffs = []
        for q in range(j * K, (j + 1) * K):
            diff = Z[:, q][:, None] - Z[:, q][None, :]
            diffs.append(diff)
            sq += diff * diff
        e = np.exp(-0.5 * sq)
        dc[j] = (G * e).sum()
        w = G * e * cc[j]
        for mm, q in enumerate(range(j * K, (j + 1) * K)):
            wd = w * diffs[mm]
            dZ[:, q] = -wd.sum(axis=1) + wd.sum(axis=0)
    dnoise = np.trace(G)
    dmean = alpha.sum()
    return dZ, dc, dnoise, dmean


# ----------------------------------------------------------------------------------------------
# exact GP prediction (training_routines.py:539-585)
# ----------------------------------------------------------------------------------------------
def predict_dense(Ztrain, Ztest, c, J, K, noise, y, mean_const=0.0, full_cov=False):
    Kxx = additive_rbf_dense(Ztrain, Ztrain, c, J, K) + noise * np.eye(Ztrain.shape[0])
    Ksx = additive_rbf_dense(Ztest, Ztrain, c, J, K)
    Lc = np.linalg.cholesky(Kxx)
    alpha = np.linalg.solve(Lc.T, np.linalg.solve(Lc, np.asarray(y, dtype=np.float64) - mean_const))
    mean = Ksx @ alpha + mean_const
    v = np.linalg.solve(Lc, Ksx.T)
    if full_cov:
        cov = additive_rbf_dense(Ztest, Ztest, c, J, K) - v.T @ v
        return mean, cov
    var = kernel_diag(c, J, Ztest.shape[0]) - (v * v).sum(axis=0)
    return mean, var
