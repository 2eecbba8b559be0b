"""CPU tests of the host-side solver (CG, Lanczos/SLQ, pivoted Cholesky, inv_quad_logdet autograd) on an explicit
matrix (DenseLazyTensor), against dense linear algebra.  The K.V operator itself is GPU-only and covered by -m gpu tests.
"""
import math
import warnings

import numpy as np
import pytest
import torch

from oracle import rpgp_oracle as orc
from rpgp.gp import settings
from rpgp.lazy import AddedDiagLazyTensor, DenseLazyTensor
from rpgp.solver import PivCholPreconditioner, linear_cg, pivoted_cholesky, slq_logdet


def make_K(n, J=5, seed=0, dtype=torch.float64):
    rng = np.random.RandomState(seed)
    Z = rng.randn(n, J)
    c = rng.rand(J) + 0.2
    return torch.from_numpy(orc.additive_rbf_dense(Z, Z, c, J, 1)).to(dtype)


def test_linear_cg_matches_dense_solve_and_masks_zero_columns():
    n = 200
    K = make_K(n) + 2.0 * torch.eye(n, dtype=torch.float64)
    rhs = torch.randn(n, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    rhs[:, 2] = 0
    x, info = linear_cg(K.matmul, rhs, tolerance=1e-9, max_iter=1000, return_info=True)
    np.testing.assert_allclose(x.numpy(), torch.linalg.solve(K, rhs).numpy(), atol=1e-7)
    assert torch.all(x[:, 2] == 0) and info["converged"]
    # at least 11 iterations are always run (k >= 10 before the tolerance test), like GPyTorch
    _, info2 = linear_cg(K.matmul, rhs, tolerance=1e3, max_iter=1000, return_info=True)
    assert info2["iterations"] == 11


def test_linear_cg_warns_when_not_converged():
    n = 120
    K = make_K(n) + 1e-3 * torch.eye(n, dtype=torch.float64)
    rhs = torch.randn(n, 2, dtype=torch.float64)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        linear_cg(K.matmul, rhs, tolerance=1e-14, max_iter=3)
    assert len(w) == 1 and "CG terminated" in str(w[0].message)


def test_lanczos_tridiagonal_gives_logdet():
    n = 150
    K = make_K(n, seed=1) + 3.0 * torch.eye(n, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    probes = torch.randn(n, 200, dtype=torch.float64, generator=g)
    probes = probes / probes.norm(dim=0, keepdim=True)
    _, T = linear_cg(K.matmul, probes, n_tridiag=200, tolerance=1e-10, max_iter=1000, max_tridiag_iter=60)
    assert T.shape[0] == 200 and T.shape[1] == T.shape[2]
    est = float(slq_logdet(T, n))
    exact = float(torch.logdet(K))
    assert abs(est - exact) / abs(exact) < 0.03, (est, exact)


def test_pivoted_cholesky_and_preconditioner_algebra():
    n = 300
    Kmat = make_K(n, J=3, seed=2)
    op = DenseLazyTensor(Kmat)
    L = pivoted_cholesky(op, 15)
    assert L.shape == (n, 15)
    # rank-15 approximation error decreases and the factor reproduces the pivot rows exactly
    err = (Kmat - L @ L.t()).diagonal()
    assert float(err.min()) > -1e-8 and float(err.sum()) < float(Kmat.diagonal().sum())
    noise = torch.tensor(0.3, dtype=torch.float64)
    P = PivCholPreconditioner(L, noise)
    Pd = L @ L.t() + noise * torch.eye(n, dtype=torch.float64)
    v = torch.randn(n, 3, dtype=torch.float64)
    np.testing.assert_allclose(P.solve(v).numpy(), torch.linalg.solve(Pd, v).numpy(), atol=1e-9)
    np.testing.assert_allclose(float(P.logdet), float(torch.logdet(Pd)), rtol=1e-10)
    s = P.sample(20000, generator=torch.Generator().manual_seed(0))
    emp = (s @ s.t()) / s.shape[1]
    assert float((emp - Pd).abs().max()) < 0.15
    # preconditioned CG converges in fewer iterations than plain CG
    Khat = Kmat + noise * torch.eye(n, dtype=torch.float64)
    rhs = torch.randn(n, 2, dtype=torch.float64)
    _, plain = linear_cg(Khat.matmul, rhs, tolerance=1e-8, max_iter=2000, return_info=True)
    xp, pre = linear_cg(Khat.matmul, rhs, tolerance=1e-8, max_iter=2000, preconditioner=P.solve, return_info=True)
    assert pre["iterations"] <= plain["iterations"]
    np.testing.assert_allclose(xp.numpy(), torch.linalg.solve(Khat, rhs).numpy(), atol=1e-6)


@pytest.mark.parametrize("n,precond", [(150, False), (260, True)])
def test_inv_quad_logdet_cg_path_values_and_gradients(n, precond):
    Kmat = make_K(n, seed=4).requires_grad_(True)
    noise = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)
    y = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).requires_grad_(True)
    # exact
    Khat = Kmat + noise * torch.eye(n, dtype=torch.float64)
    iq_ref = y @ torch.linalg.solve(Khat, y)
    ld_ref = torch.logdet(Khat)
    g_ref = torch.autograd.grad(iq_ref + ld_ref, [Kmat, noise, y])
    giq_ref = torch.autograd.grad(y @ torch.linalg.solve(Kmat + noise * torch.eye(n, dtype=torch.float64), y), [Kmat, noise, y])
    with settings.max_cholesky_size(0), settings.cg_tolerance(1e-9), settings.max_cg_iterations(4000), \
            settings.num_trace_samples(400), settings.max_lanczos_quadrature_iterations(80), \
            settings.min_preconditioning_size(0 if precond else 10 ** 9):
        torch.manual_seed(0)
        op = AddedDiagLazyTensor(DenseLazyTensor(Kmat), noise)
        iq, ld = op.inv_quad_logdet(inv_quad_rhs=y.unsqueeze(-1), logdet=True)
        np.testing.assert_allclose(float(iq), float(iq_ref), rtol=1e-7)
        assert abs(float(ld) - float(ld_ref)) / abs(float(ld_ref)) < 0.05
        giq = torch.autograd.grad(iq, [Kmat, noise, y], retain_graph=True)
        for a, b in zip(giq, giq_ref):   # the inverse-quadratic part is deterministic: exact to CG tolerance
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-7)
        g = torch.autograd.grad(iq + ld, [Kmat, noise, y])
        # logdet gradient is a Hutchinson estimate: tr(K^-1 dK) with 400 probes
        assert abs(float(g[1]) - float(g_ref[1])) / abs(float(g_ref[1])) < 0.1
        np.testing.assert_allclose(g[2].numpy(), g_ref[2].numpy(), rtol=1e-5, atol=1e-7)


def test_inv_quad_logdet_cholesky_path_is_exact():
    n = 90
    Kmat = make_K(n, seed=5).requires_grad_(True)
    noise = torch.tensor(0.2, dtype=torch.float64, requires_grad=True)
    y = torch.randn(n, dtype=torch.float64)
    op = AddedDiagLazyTensor(DenseLazyTensor(Kmat), noise)
    iq, ld = op.inv_quad_logdet(inv_quad_rhs=y.unsqueeze(-1), logdet=True)   # n <= max_cholesky_size
    Khat = Kmat.detach() + noise.detach() * torch.eye(n, dtype=torch.float64)
    mll, alpha, logdet, inv_quad = orc.exact_mll_dense(Khat.numpy(), y.numpy())
    np.testing.assert_allclose(float(iq), inv_quad, rtol=1e-10)
    np.testing.assert_allclose(float(ld), logdet, rtol=1e-10)
    (iq + ld).backward()
    Kinv = np.linalg.inv(Khat.numpy())
    np.testing.assert_allclose(float(noise.grad), np.trace(Kinv) - alpha @ alpha, rtol=1e-8)


def test_psd_safe_cholesky_retries_with_jitter_like_gpytorch():
    from rpgp.solver.inv_quad_logdet import psd_safe_cholesky
    g = torch.Generator().manual_seed(0)
    B = torch.randn(40, 6, generator=g)
    A = B @ B.T                                              # rank 6: plain Cholesky fails
    with pytest.raises(RuntimeError):
        torch.linalg.cholesky(A - 1e-7 * torch.eye(40))
    L = psd_safe_cholesky(A - 1e-7 * torch.eye(40))
    assert torch.isfinite(L).all() and float((L @ L.T - A).abs().max()) < 1e-3
    good = A + torch.eye(40)
    assert torch.equal(psd_safe_cholesky(good), torch.linalg.cholesky(good))      # untouched when the matrix factors
    with pytest.raises(RuntimeError):
        psd_safe_cholesky(A - 10.0 * torch.eye(40))           # far from positive definite: the original error comes back
    bad = A.clone()
    bad[0, 0] = float("nan")
    with pytest.raises(RuntimeError):
        psd_safe_cholesky(bad)


def test_settings_context_managers_nest_and_restore():
    assert settings.cg_tolerance.value() == 1.0 and settings.eval_cg_tolerance.value() == 0.01
    assert settings.max_cg_iterations.value() == 1000 and settings.max_cholesky_size.value() == 800
    assert settings.max_preconditioner_size.value() == 15 and settings.min_preconditioning_size.value() == 2000
    assert settings.num_trace_samples.value() == 10
    with settings.cg_tolerance(0.002), settings.eval_cg_tolerance(0.001), settings.max_cg_iterations(10_000):
        assert settings.cg_tolerance.value() == 0.002 and settings.max_cg_iterations.value() == 10_000
        with settings.fast_computations(False, False, False):
            assert settings.fast_computations.solves.off() and settings.fast_computations.log_prob.off()
        assert settings.fast_computations.solves.on()
    assert settings.cg_tolerance.value() == 1.0
    with settings.skip_posterior_variances(True):
        assert settings.skip_posterior_variances.on()
    assert settings.skip_posterior_variances.off()
