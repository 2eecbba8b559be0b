"""CPU tests of the lazy predictive covariance (rpgp/lazy.py PredictiveCovarLazyTensor) and the Lanczos root behind
settings.fast_pred_var (rpgp/solver/lanczos.py), with explicit matrices standing in for the fused operators -- the GPU tests
(tests/test_model_gpu.py) run the same algebra over the CUDA kernels."""
import numpy as np
import torch

from rpgp import lazy
from rpgp.gp import settings
from rpgp.solver.lanczos import lanczos_root_inv, lanczos_tridiag


def _problem(n=60, m=25, seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, 3, generator=g, dtype=torch.float64)
    Xs = torch.randn(m, 3, generator=g, dtype=torch.float64)

    def k(a, b):
        return torch.exp(-0.5 * torch.cdist(a, b) ** 2)

    return k(X, X), k(Xs, X), k(Xs, Xs), 0.3


def test_lanczos_tridiagonalises_and_inverts():
    K, _, _, noise = _problem()
    A = K + noise * torch.eye(K.shape[0], dtype=K.dtype)
    v = torch.ones(K.shape[0], dtype=K.dtype)
    Q, T = lanczos_tridiag(lambda x: A @ x, v, 60)
    k = Q.shape[1]
    np.testing.assert_allclose((Q.t() @ Q).numpy(), np.eye(k), atol=1e-8)
    np.testing.assert_allclose((Q.t() @ A @ Q).numpy(), T.numpy(), atol=1e-8)
    W = lanczos_root_inv(lambda x: A @ x, v, 60)
    if W.shape[1] == 60:      # full Krylov space: W W^T is the inverse
        np.testing.assert_allclose((W @ W.t()).numpy(), torch.linalg.inv(A).numpy(), atol=1e-6)
    # in any case A^-1 v is reproduced (v spans the first Krylov vector)
    np.testing.assert_allclose((W @ (W.t() @ v)).numpy(), torch.linalg.solve(A, v).numpy(), rtol=1e-6, atol=1e-8)


def test_lazy_predictive_covariance_matches_the_dense_formula():
    K, Ks, Kss, noise = _problem()
    n = K.shape[0]
    ref = Kss - Ks @ torch.linalg.solve(K + noise * torch.eye(n, dtype=K.dtype), Ks.t())
    train = lazy.AddedDiagLazyTensor(lazy.DenseLazyTensor(K), torch.tensor(noise, dtype=K.dtype))
    with settings.variance_batch_size(7):
        cov = lazy.PredictiveCovarLazyTensor(lazy.DenseLazyTensor(Kss), lazy.DenseLazyTensor(Ks), train)
        np.testing.assert_allclose(cov.diag().numpy(), ref.diagonal().numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(cov.evaluate().numpy(), ref.numpy(), rtol=1e-9, atol=1e-12)
    V = torch.randn(Ks.shape[0], 3, dtype=K.dtype, generator=torch.Generator().manual_seed(1))
    np.testing.assert_allclose(cov._matmul(V).numpy(), (ref @ V).numpy(), rtol=1e-9, atol=1e-12)
    idx = torch.tensor([3, 0, 11])
    np.testing.assert_allclose(cov.rows(idx).numpy(), ref[idx].numpy(), rtol=1e-9, atol=1e-12)
    assert cov.shape == (Ks.shape[0], Ks.shape[0])


def test_love_root_reproduces_the_exact_covariance_at_full_rank_and_approximates_below():
    K, Ks, Kss, noise = _problem(n=80, m=30, seed=2)
    n = K.shape[0]
    A = K + noise * torch.eye(n, dtype=K.dtype)
    ref = Kss - Ks @ torch.linalg.solve(A, Ks.t())
    train = lazy.AddedDiagLazyTensor(lazy.DenseLazyTensor(K), torch.tensor(noise, dtype=K.dtype))
    init = Ks.t().mean(dim=1)
    errs = []
    for rank in (10, 30, 80):
        W = lanczos_root_inv(lambda x: A @ x, init, rank)
        cov = lazy.PredictiveCovarLazyTensor(lazy.DenseLazyTensor(Kss), lazy.DenseLazyTensor(Ks), train, root=W)
        errs.append(float((cov.diag() - ref.diagonal()).abs().max()))
        np.testing.assert_allclose(cov.evaluate().numpy(), (Kss - Ks @ W @ W.t() @ Ks.t()).numpy(), atol=1e-10)
    assert errs[2] < 1e-6 and errs[1] < errs[0] and errs[1] < 0.2      # monotone in the rank, exact at full rank


def test_joint_log_prob_through_the_lazy_covariance():
    """test_nll of train_exact_gp (training_routines.py:567): log N(y* | mean, Sigma* + sigma^2 I) with Sigma* lazy"""
    from rpgp.gp.distributions import MultivariateNormal
    K, Ks, Kss, noise = _problem(n=50, m=20, seed=3)
    n, m = K.shape[0], Ks.shape[0]
    ref = Kss - Ks @ torch.linalg.solve(K + noise * torch.eye(n, dtype=K.dtype), Ks.t())
    train = lazy.AddedDiagLazyTensor(lazy.DenseLazyTensor(K), torch.tensor(noise, dtype=K.dtype))
    cov = lazy.PredictiveCovarLazyTensor(lazy.DenseLazyTensor(Kss), lazy.DenseLazyTensor(Ks), train)
    mean = torch.zeros(m, dtype=K.dtype)
    ys = torch.randn(m, dtype=K.dtype, generator=torch.Generator().manual_seed(4))
    dist = MultivariateNormal(mean, cov.add_diag(torch.tensor(noise, dtype=K.dtype)))
    want = torch.distributions.MultivariateNormal(mean, ref + noise * torch.eye(m, dtype=K.dtype)).log_prob(ys)
    np.testing.assert_allclose(float(dist.log_prob(ys)), float(want), rtol=1e-9)
    np.testing.assert_allclose(dist.variance.numpy(), (ref.diagonal() + noise).numpy(), rtol=1e-9)


def test_joint_log_prob_through_cg_on_the_lazy_covariance():
    """large n*: the joint test log-probability runs CG + stochastic Lanczos quadrature (with the pivoted-Cholesky preconditioner,
    which fetches rows and the diagonal of the LAZY predictive covariance) instead of a dense Cholesky factor"""
    from rpgp.gp.distributions import MultivariateNormal
    K, Ks, Kss, noise = _problem(n=70, m=60, seed=5)
    n, m = K.shape[0], Ks.shape[0]
    ref = Kss - Ks @ torch.linalg.solve(K + noise * torch.eye(n, dtype=K.dtype), Ks.t())
    train = lazy.AddedDiagLazyTensor(lazy.DenseLazyTensor(K), torch.tensor(noise, dtype=K.dtype))
    cov = lazy.PredictiveCovarLazyTensor(lazy.DenseLazyTensor(Kss), lazy.DenseLazyTensor(Ks), train)
    ys = torch.randn(m, dtype=K.dtype, generator=torch.Generator().manual_seed(6))
    full = ref + noise * torch.eye(m, dtype=K.dtype)
    want_iq = float(ys @ torch.linalg.solve(full, ys))
    want_ld = float(torch.logdet(full))
    probes = torch.randn(m, 200, dtype=K.dtype, generator=torch.Generator().manual_seed(7))
    for precond_size in (0, 10):
        op = cov.add_diag(torch.tensor(noise, dtype=K.dtype))
        with settings.max_cholesky_size(0), settings.min_preconditioning_size(1), settings.max_preconditioner_size(precond_size), \
                settings.cg_tolerance(1e-10), settings.eval_cg_tolerance(1e-12), settings.num_trace_samples(200), \
                settings.deterministic_probes(probes):      # (the inner solves with K + sigma^2 I are CG too: nested)
            assert not op._use_cholesky()
            iq, ld = op.inv_quad_logdet(inv_quad_rhs=ys.unsqueeze(-1), logdet=True)
            assert (op._preconditioner() is not None) == (precond_size > 0)
        np.testing.assert_allclose(float(iq), want_iq, rtol=1e-7)
        assert abs(float(ld) - want_ld) < 0.05 * abs(want_ld) + 0.5, (float(ld), want_ld)      # SLQ: 200 probes
        dist = MultivariateNormal(torch.zeros(m, dtype=K.dtype), op)
        assert np.isfinite(float(dist.variance.min()))
