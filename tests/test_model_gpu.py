"""GPU parity of the reference-facing API (kernel classes -> fused operator -> solver) against the CPU oracle.

Golden vectors G1-G6 of SURVEY.md §8c are asserted through the same classes the reference's test.py uses; MLL values,
gradients and predictions -- which no reference test pins -- are checked against the dense FP64 Cholesky oracle with
the tolerances of BASELINE.json (MLL / gradients 1e-4 relative on the exact path; trained-model predictions 1 %).
"""
import copy
import json
import math
import os
import warnings

import numpy as np
import pytest
import torch

import training_routines as tr
from fitting.optimizing import mean_squared_error, train_to_convergence
from gp_models.kernels import (GAMFunction, MemoryEfficientGamKernel, PolynomialProjectionKernel, ScaledProjectionKernel)
from gp_models.models import ExactGPModel
from oracle import rpgp_oracle as orc
from rpgp import gp as gpytorch
from rpgp.gp import settings
from rpgp.gp.kernels import AdditiveStructureKernel, RBFKernel, ScaleKernel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = torch.device("cuda:0")


from parity_util import rel  # noqa: E402  (norm-wise error; its repr adds the row-wise / element-wise figures)


# ---- golden vectors through the reference-facing classes ---------------------------------------------------------------
@pytest.mark.parametrize("prescale", [True, False])
def test_g1_g2_manual_rescale_kernel(prescale):
    x = torch.tensor([[1., 2., 3.], [1.1, 2.2, 3.3]], device=DEV)
    kbase = RBFKernel()
    kbase.initialize(lengthscale=torch.tensor([1.]))
    proj = torch.nn.Linear(3, 3, bias=False)
    proj.weight.data = torch.eye(3)
    k = ScaledProjectionKernel(proj, AdditiveStructureKernel(kbase, 3), prescale=prescale, ard_num_dims=3)
    k.initialize(lengthscale=torch.tensor([1., 2., 3.]))
    k = k.to(DEV)
    with torch.no_grad():
        K = k(x, x).evaluate().cpu().numpy()
    np.testing.assert_allclose(K, [[3.0, 2.985037326813], [2.985037326813, 3.0]], rtol=1e-6)
    k1 = RBFKernel().to(DEV)
    k1.initialize(lengthscale=torch.tensor([1.]))
    with torch.no_grad():
        K2 = 3 * k1(x[:, 0:1], x[:, 0:1]).evaluate().cpu().numpy()
    np.testing.assert_allclose(K, K2, rtol=1e-6)


def test_g3_memory_efficient_gam_equals_additive_structure():
    x = torch.tensor([[1., 2., 3.], [1.1, 2.2, 3.3]], device=DEV)
    K = MemoryEfficientGamKernel().to(DEV)(x, x).evaluate().detach().cpu().numpy()
    k = ScaleKernel(RBFKernel())
    k.initialize(outputscale=1.)
    K2 = AdditiveStructureKernel(k, 2).to(DEV)(x, x).evaluate().detach().cpu().numpy()
    np.testing.assert_allclose(K, K2, atol=1e-6)
    np.testing.assert_allclose(K, [[3.0, 2.859465122223], [2.859465122223, 3.0]], atol=1e-6)


def test_g4_g5_gamfunction_forward_backward_gradcheck():
    g = np.load(os.path.join(GOLD, "gam_g4_g5.npz"))
    x1 = torch.tensor(g["x1"], device=DEV, requires_grad=True)
    x2 = torch.tensor(g["x2"], device=DEV, requires_grad=True)
    raw = torch.tensor(g["raw_lengthscale"], device=DEV, requires_grad=True)
    ls = torch.nn.functional.softplus(raw)
    K = GAMFunction.apply(x1, x2, ls)
    np.testing.assert_allclose(K.detach().cpu().numpy(), g["K"], rtol=1e-12)
    K.sum().backward()
    np.testing.assert_allclose(x1.grad.cpu().numpy(), g["dx1"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(x2.grad.cpu().numpy(), g["dx2"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(raw.grad.cpu().numpy(), g["draw"], rtol=1e-10)
    a = torch.tensor(g["x1"], device=DEV, requires_grad=True)
    b = torch.tensor(g["x2"], device=DEV, requires_grad=True)
    l2 = torch.nn.functional.softplus(torch.tensor(g["raw_lengthscale"], device=DEV)).requires_grad_(True)
    assert torch.autograd.gradcheck(GAMFunction.apply, (a, b, l2))
    with pytest.raises(ValueError, match="Dimension mismatch"):
        GAMFunction.apply(a, b[:, :2], l2)


@pytest.mark.parametrize("idx", [0, 2, 3])
def test_gamfunction_random_fixtures_fp32(idx):
    g = np.load(os.path.join(GOLD, "gam_random.npz"))
    pre = "c%d_f32_" % idx
    x1 = torch.tensor(g[pre + "x1"], device=DEV, requires_grad=True)
    x2 = torch.tensor(g[pre + "x2"], device=DEV, requires_grad=True)
    ell = torch.tensor(g[pre + "ell"], device=DEV, requires_grad=True)
    S = torch.tensor(g[pre + "L"] @ g[pre + "R"].T, device=DEV)
    K = GAMFunction.apply(x1, x2, ell)
    assert rel(K.detach().cpu().numpy(), g["c%d_f64_K" % idx]) < 1e-5
    (K * S).sum().backward()
    assert rel(x1.grad.cpu().numpy(), g["c%d_f64_dx1" % idx]) < 1e-4
    assert rel(x2.grad.cpu().numpy(), g["c%d_f64_dx2" % idx]) < 1e-4
    assert rel(ell.grad.cpu().numpy(), g["c%d_f64_dell" % idx]) < 1e-4


def test_g6_polynomial_projection_forward():
    torch.manual_seed(0)
    d, J, k = 7, 3, 2
    x = torch.randn(40, d, device=DEV)
    kernel = PolynomialProjectionKernel(J, k, d, RBFKernel, [torch.eye(d, k)] * J, [torch.zeros(k)] * J).to(DEV)
    out = kernel(x, x)
    assert isinstance(out, gpytorch.kernels.LazyEvaluatedKernelTensor)
    K_proj = out.evaluate().detach().cpu().numpy()
    k1 = RBFKernel().to(DEV)
    Keq = (k1(x[:, :1]).evaluate() * k1(x[:, 1:2]).evaluate()).detach().cpu().numpy()   # (1/3) * 3 * k(x0) k(x1)
    np.testing.assert_allclose(K_proj, Keq, atol=1e-5)


# ---- dense FP64 reference of a whole model (same parameters, plain torch on CPU) -----------------------------------------
def dense_reference_loss(model, X, y):
    """-MLL/n of `model` evaluated densely in FP64 on the CPU from the model's own lowering (Z, c) -- differentiable."""
    m = copy.deepcopy(model).to("cpu", torch.float64)
    m.train()
    Xc, yc = X.detach().to("cpu", torch.float64), y.detach().to("cpu", torch.float64)
    op = m.covar_module(Xc).evaluate_kernel()
    Z, c, J, K = op.Z1, op.c, op.J, op.K
    n = Z.shape[0]
    Zg = Z.reshape(n, J, K)
    sq = ((Zg[:, None] - Zg[None]) ** 2).sum(-1)                  # (n, n, J)
    Kd = (torch.exp(-0.5 * sq) * c).sum(-1)
    noise = m.likelihood.noise
    Khat = Kd + noise * torch.eye(n, dtype=torch.float64)
    r = yc - m.mean_module(Xc)
    Lc = torch.linalg.cholesky(Khat)
    sol = torch.linalg.solve_triangular(Lc, r.unsqueeze(-1), upper=False)
    mll = -0.5 * ((sol ** 2).sum() + 2 * Lc.diagonal().log().sum() + n * math.log(2 * math.pi))
    for _, prior, closure in gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m).named_priors():
        mll = mll + prior.log_prob(closure()).sum()
    return -mll / n, m


def synthetic(n, d, seed, device=DEV, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, generator=g) * 4 - 2
    y = torch.sin(X).sum(-1) + 0.05 * torch.randn(n, generator=g)
    y = (y - y.mean()) / y.std()
    return X.to(device, dtype), y.to(device, dtype)


def build(spec, X, y, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    np.random.seed(seed)
    s = tr.load_model_spec(spec)
    kw = {k: (X.shape[1] if v == "d" else v) for k, v in s["model_kwargs"].items()}
    model, lik = tr.create_exact_gp(X, y, s["kind"], **kw)
    model = model.to(X.device, dtype)
    return model, lik, gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)


def perturb(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.add_((0.3 * torch.randn(p.shape, generator=g)).to(p.device, p.dtype))


@pytest.mark.parametrize("spec,dtype,tol", [
    ("additive_rp_J20_K1", torch.float32, 1e-4), ("additive_rp_prescale_J20", torch.float32, 1e-4),
    ("additive_rp_postscale_J20", torch.float32, 1e-4), ("additive_spread_prescale_Jd", torch.float32, 1e-4),
    ("additive_rp_prescale_J1_K20", torch.float32, 1e-4), ("additive_rp_prescale_J20_K5", torch.float32, 1e-4),
    ("additive_rp_prescale_J20", torch.float64, 1e-8), ("additive_rp_J20_K1", torch.float64, 1e-8),
])
def test_exact_mll_and_gradients_match_dense_oracle(spec, dtype, tol):
    X, y = synthetic(300, 6, seed=1, dtype=dtype)
    model, lik, mll = build(spec, X, y, dtype=dtype)
    perturb(model)
    model.train()
    loss = -mll(model(X), y)                                       # n = 300 <= max_cholesky_size: exact path
    loss.backward()
    ref_loss, ref_model = dense_reference_loss(model, X, y)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < tol
    ref_grads = dict(ref_model.named_parameters())
    checked = 0
    for name, p in model.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        want = ref_grads[name].grad.numpy()
        scale = max(np.abs(want).max(), 1e-3)
        assert np.abs(p.grad.cpu().numpy() - want).max() / scale < tol * 10, name
        assert rel(p.grad.cpu().numpy(), want) < tol * 10, name
        checked += 1
    assert checked >= 4
    # cross-check the oracle's closed-form gradient of n*MLL w.r.t. the noise (dense numpy path)
    op = ref_model.covar_module(X.detach().to("cpu", torch.float64)).evaluate_kernel()
    _, _, dnoise, _ = orc.exact_mll_grads_dense(op.Z1.detach().numpy(), op.c.detach().numpy(), op.J, op.K,
                                                ref_model.likelihood.noise.item(), y.cpu().numpy().astype(np.float64),
                                                ref_model.mean_module.constant.item())
    assert np.isfinite(dnoise)


def test_cg_path_mll_matches_dense_oracle_within_estimator_noise():
    X, y = synthetic(2500, 8, seed=2)                               # > max_cholesky_size, >= min_preconditioning_size
    model, lik, mll = build("additive_rp_prescale_J20", X, y)
    model.train()
    torch.manual_seed(0)
    with settings.cg_tolerance(1e-4), settings.max_cg_iterations(2000), settings.num_trace_samples(100), \
            warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = -mll(model(X), y)
        loss.backward()
    ref_loss, ref_model = dense_reference_loss(model, X, y)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < 0.01
    ref_grads = dict(ref_model.named_parameters())
    for name, p in model.named_parameters():
        if p.requires_grad:
            want = ref_grads[name].grad.numpy()
            got = p.grad.cpu().numpy()
            assert np.abs(got - want).max() < 0.1 * max(np.abs(want).max(), 0.02), (name, got, want)


def test_cg_inverse_quadratic_term_is_exact():
    """the deterministic half of the CG path: y^T K^-1 y and its gradient, 1e-4 relative"""
    X, y = synthetic(1500, 8, seed=3)
    model, lik, mll = build("additive_rp_prescale_J20", X, y)
    model.train()
    with settings.cg_tolerance(1e-5), settings.max_cg_iterations(3000), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        covar = lik(model(X)).lazy_covariance_matrix
        iq, _ = covar.inv_quad_logdet(inv_quad_rhs=(y - model.mean_module(X)).unsqueeze(-1), logdet=False)
        iq.backward()
    m = copy.deepcopy(model).to("cpu", torch.float64)
    for p in m.parameters():
        p.grad = None
    Xc, yc = X.cpu().double(), y.cpu().double()
    op = m.covar_module(Xc).evaluate_kernel()
    Zg = op.Z1.reshape(1500, op.J, op.K)
    Kd = (torch.exp(-0.5 * ((Zg[:, None] - Zg[None]) ** 2).sum(-1)) * op.c).sum(-1)
    Khat = Kd + m.likelihood.noise * torch.eye(1500, dtype=torch.float64)
    r = yc - m.mean_module(Xc)
    iq_ref = r @ torch.linalg.solve(Khat, r)
    iq_ref.backward()
    assert abs(iq.item() - iq_ref.item()) / iq_ref.item() < 1e-4
    ref = dict(m.named_parameters())
    for name, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:
            assert rel(p.grad.cpu().numpy(), ref[name].grad.numpy()) < 2e-3, name


@pytest.mark.parametrize("n,tol", [(400, 1e-4), (2200, 2e-3)])
def test_prediction_matches_dense_oracle(n, tol):
    X, y = synthetic(n, 6, seed=4)
    Xt, yt = synthetic(150, 6, seed=5)
    model, lik, mll = build("additive_rp_J20_K1", X, y)
    perturb(model, 1)
    model.eval()
    lik.eval()
    with torch.no_grad(), settings.eval_cg_tolerance(1e-5), settings.max_cg_iterations(3000), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(Xt)
        mean, var = out.mean.cpu().numpy(), out.variance.cpu().numpy()
        lower, upper = lik(out).confidence_region()
    m = copy.deepcopy(model).to("cpu", torch.float64)
    opx = m.covar_module(X.cpu().double()).evaluate_kernel()
    opt = m.covar_module(Xt.cpu().double(), X.cpu().double()).evaluate_kernel()
    ref_mean, ref_var = orc.predict_dense(opx.Z1.detach().numpy(), opt.Z1.detach().numpy(), opx.c.detach().numpy(), opx.J,
                                          opx.K, m.likelihood.noise.item(), y.cpu().numpy().astype(np.float64),
                                          m.mean_module.constant.item())
    assert rel(mean, ref_mean) < tol
    assert rel(var, ref_var) < max(tol, 1e-3)
    assert torch.all(upper > lower)
    with torch.no_grad(), settings.skip_posterior_variances(True), settings.eval_cg_tolerance(1e-5):
        out2 = model(Xt)
    assert rel(out2.mean.cpu().numpy(), ref_mean) < tol and float(out2.variance.max()) <= 1e-9


def _dense_prediction(model, X, Xt, y):
    m = copy.deepcopy(model).to("cpu", torch.float64)
    opx = m.covar_module(X.cpu().double()).evaluate_kernel()
    opt = m.covar_module(Xt.cpu().double(), X.cpu().double()).evaluate_kernel()
    return orc.predict_dense(opx.Z1.detach().numpy(), opt.Z1.detach().numpy(), opx.c.detach().numpy(), opx.J, opx.K,
                             m.likelihood.noise.item(), y.cpu().numpy().astype(np.float64), m.mean_module.constant.item())


def test_lazy_predictive_variances_match_dense_oracle():
    """predictive variances without the n* x n cross-covariance -- batches of test points, one multi-right-hand-side
    CG solve each (through the symmetric tensor-core kernel), against the dense FP64 oracle; then the joint test log-probability
    (`test_nll` of train_exact_gp, training_routines.py:567) through the same lazy covariance."""
    n, nt = 1500, 130
    X, y = synthetic(n, 6, seed=14)
    Xt, yt = synthetic(nt, 6, seed=15)
    model, lik, mll = build("additive_rp_J20_K1", X, y)
    perturb(model, 2)
    model.eval()
    lik.eval()
    ref_mean, ref_var = _dense_prediction(model, X, Xt, y)
    with torch.no_grad(), settings.eval_cg_tolerance(1e-6), settings.max_cg_iterations(4000), settings.max_dense_predictive_size(0), \
            settings.variance_batch_size(48), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(Xt)
        from rpgp.lazy import PredictiveCovarLazyTensor
        assert isinstance(out.lazy_covariance_matrix, PredictiveCovarLazyTensor)
        var = out.variance.cpu().numpy()
        lazy_dense = out.lazy_covariance_matrix.evaluate().cpu().numpy()
        nll_lazy = -float(lik(out).log_prob(yt)) / nt
    with torch.no_grad(), settings.eval_cg_tolerance(1e-6), settings.max_cg_iterations(4000), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out_dense = model(Xt)
        nll_dense = -float(lik(out_dense).log_prob(yt)) / nt
    assert rel(out.mean.cpu().numpy(), ref_mean) < 2e-3
    assert rel(var, ref_var) < 2e-3, rel(var, ref_var)
    assert rel(np.diag(lazy_dense), ref_var) < 2e-3
    assert rel(lazy_dense, out_dense.covariance_matrix.cpu().numpy()) < 2e-3
    assert abs(nll_lazy - nll_dense) < 1e-3 * max(1.0, abs(nll_dense)), (nll_lazy, nll_dense)


def test_love_fast_predictive_variances():
    """settings.fast_pred_var (gp_experiment_runner.py:235,327): cached Lanczos root of K^-1.  At full rank it reproduces the exact
    variances; at the default-style low rank it stays an upper bound of them (the root under-estimates K^-1) within a few percent
    of the prior variance."""
    n, nt = 700, 90
    X, y = synthetic(n, 6, seed=24)
    Xt, _ = synthetic(nt, 6, seed=25)
    model, lik, mll = build("additive_rp_J20_K1", X, y)
    perturb(model, 3)
    model.eval()
    lik.eval()
    _, ref_var = _dense_prediction(model, X, Xt, y)
    prior_var = float(model.covar_module(Xt).evaluate_kernel().diag().detach().max())
    with torch.no_grad(), settings.fast_pred_var(True), settings.max_root_decomposition_size(n), settings.eval_cg_tolerance(1e-6), \
            warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = model(Xt).variance.cpu().numpy()
    assert rel(full, ref_var) < 5e-3, rel(full, ref_var)
    model.train()
    model.eval()
    with torch.no_grad(), settings.fast_pred_var(True), settings.max_root_decomposition_size(100), settings.eval_cg_tolerance(1e-6), \
            warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(Xt)
        low = out.variance.cpu().numpy()
        assert out.lazy_covariance_matrix.root.shape[1] <= 100
    assert np.all(low > ref_var - 1e-3 * prior_var)
    assert np.abs(low - ref_var).max() < 0.25 * prior_var


def test_one_adam_step_moves_the_right_parameters():
    # test.py:575-621: lengthscale moves, frozen base lengthscale and W stay; with learn_proj W moves too
    x = torch.tensor([[1., 2., 3.], [1.1, 2.2, 3.3]], device=DEV)
    y = torch.sin(x).sum(dim=1)
    for learn in (False, True):
        kbase = RBFKernel()
        kbase.initialize(lengthscale=torch.tensor([1.]))
        proj = torch.nn.Linear(3, 3, bias=False)
        proj.weight.data = torch.eye(3)
        pk = ScaledProjectionKernel(proj, AdditiveStructureKernel(kbase, 3), prescale=True, ard_num_dims=3, learn_proj=learn)
        pk.initialize(lengthscale=torch.tensor([1., 2., 3.]))
        model = ExactGPModel(x, y, gpytorch.likelihoods.GaussianLikelihood(), pk).to(DEV)
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(model.likelihood, model)
        opt = torch.optim.Adam(model.parameters(), lr=0.1)
        opt.zero_grad()
        loss = -mll(model(x), y)
        loss.backward()
        opt.step()
        np.testing.assert_allclose(pk.base_kernel.base_kernel.lengthscale.detach().cpu().numpy(), [[1.]], rtol=1e-6)
        moved_W = not np.allclose(pk.projection_module.weight.detach().cpu().numpy(), np.eye(3))
        assert moved_W == learn
        assert not np.allclose(pk.lengthscale.detach().cpu().numpy(), [[1., 2., 3.]])


def test_train_exact_gp_end_to_end_cfg1_shape():
    """cfg-1 style run (synthetic_test_script.py:96-108,122-123 at reduced size): train through train_exact_gp with the
    CG path, then the trained model's held-out RMSE / NLL must agree with the dense oracle evaluated at the SAME trained
    hyper-parameters to 1 %."""
    X, y = synthetic(1200, 6, seed=6)
    Xt, yt = synthetic(400, 6, seed=7)
    spec = tr.load_model_spec("additive_rp_J20_K1")
    spec["train_kwargs"].update(max_iter=25, check_conv=False)
    torch.manual_seed(3)
    np.random.seed(3)
    with settings.cg_tolerance(0.01), settings.eval_cg_tolerance(1e-4), settings.max_cg_iterations(2000), \
            warnings.catch_warnings():
        warnings.simplefilter("ignore")
        metrics, pred_mean, model = tr.train_exact_gp(X, y, Xt, yt, spec["kind"], spec["model_kwargs"], spec["train_kwargs"],
                                                      devices=("cuda:0",), skip_random_restart=True)
    assert metrics["trained_epochs"] == 25
    rmse = float(((pred_mean - yt.cpu()) ** 2).mean().sqrt())
    assert rmse < 0.35, rmse                                          # an additive target is learnable by an additive GP
    assert metrics["prior_train_nmll"] < 1.3 and 0.5 < metrics["test_pred_frac_in_cr"] <= 1.0
    m = copy.deepcopy(model).to("cpu", torch.float64)
    opx = m.covar_module(X.cpu().double()).evaluate_kernel()
    opt = m.covar_module(Xt.cpu().double(), X.cpu().double()).evaluate_kernel()
    args = (opx.c.detach().numpy(), opx.J, opx.K, m.likelihood.noise.item(), y.cpu().numpy().astype(np.float64),
            m.mean_module.constant.item())
    ref_mean, ref_cov = orc.predict_dense(opx.Z1.detach().numpy(), opt.Z1.detach().numpy(), *args, full_cov=True)
    ref_rmse = float(np.sqrt(((ref_mean - yt.cpu().numpy()) ** 2).mean()))
    assert abs(rmse - ref_rmse) / ref_rmse < 0.01
    nt = 400
    cov = ref_cov + m.likelihood.noise.item() * np.eye(nt)
    diff = yt.cpu().numpy().astype(np.float64) - ref_mean
    Lc = np.linalg.cholesky(cov)
    sol = np.linalg.solve(Lc, diff)
    ref_nll = 0.5 * (sol @ sol + 2 * np.log(np.diag(Lc)).sum() + nt * math.log(2 * math.pi))
    ref_nll = (ref_nll - orc.smoothed_box_log_prob(m.likelihood.noise.item())) / nt
    assert abs(metrics["test_nll"] - ref_nll) / abs(ref_nll) < 0.01, (metrics["test_nll"], ref_nll)


def test_cfg1_at_its_named_size_train_and_predict():
    """BASELINE configs[0] at the size it names (synthetic_test_script.py:96-108,122-123): n = 2000 points of the 10-dimensional
    `additive` target (noise 0.01), 4000 hold-out points of the same law, both standardised by the HOLD-OUT statistics, spec
    additive_rp_J20_K1, cg_tolerance 1e-3, eval_cg_tolerance 5e-4, max_cg_iterations 10 000, Adam lr 0.1.  The trained model's
    held-out RMSE (the quantity the script reports) must agree with the dense FP64 oracle evaluated at the SAME trained
    hyper-parameters to 1 %."""
    import synthetic_test_script as sts
    g = torch.Generator().manual_seed(0)
    n, nho, d = 2000, 4000, 10
    ho_x = torch.rand(nho, d, generator=g) * 4 - 2
    ho_y = sts.additive(ho_x) + torch.randn(nho, generator=g) * 0.01
    X = torch.rand(n, d, generator=g) * 4 - 2
    y = sts.additive(X) + torch.randn(n, generator=g) * 0.01
    mx, sx, my, sy = ho_x.mean(0), ho_x.std(0), ho_y.mean(), ho_y.std()
    X, Xt, y, yt = ((X - mx) / sx).to(DEV), ((ho_x - mx) / sx).to(DEV), ((y - my) / sy).to(DEV), ((ho_y - my) / sy).to(DEV)
    spec = tr.load_model_spec("additive_rp_J20_K1")
    spec["train_kwargs"].update(max_iter=30, check_conv=False, lr=0.1)
    torch.manual_seed(0)
    np.random.seed(0)
    import time
    t0 = time.perf_counter()
    with settings.cg_tolerance(1e-3), settings.eval_cg_tolerance(5e-4), settings.max_cg_iterations(10_000), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        metrics, pred_mean, model = tr.train_exact_gp(X, y, Xt, yt, spec["kind"], spec["model_kwargs"], spec["train_kwargs"],
                                                      devices=("cuda:0",), skip_random_restart=True)
    torch.cuda.synchronize()
    print("cfg1 (n=2000, d=10, J=20): 30 epochs + evaluation in %.1f s" % (time.perf_counter() - t0))
    assert metrics["trained_epochs"] == 30
    rmse = float(((pred_mean - yt.cpu()) ** 2).mean().sqrt())
    assert rmse < 0.35, rmse                                          # (30 epochs from a random initialisation)
    m = copy.deepcopy(model).to("cpu", torch.float64)
    opx = m.covar_module(X.cpu().double()).evaluate_kernel()
    opt = m.covar_module(Xt.cpu().double(), X.cpu().double()).evaluate_kernel()
    noise = m.likelihood.noise.item()
    args = (opx.c.detach().numpy(), opx.J, opx.K, noise, y.cpu().numpy().astype(np.float64), m.mean_module.constant.item())
    ref_mean, ref_cov = orc.predict_dense(opx.Z1.detach().numpy(), opt.Z1.detach().numpy(), *args, full_cov=True)
    ref_rmse = float(np.sqrt(((ref_mean - yt.cpu().numpy()) ** 2).mean()))
    assert abs(rmse - ref_rmse) / ref_rmse < 0.01, (rmse, ref_rmse)
    # The script reports only the hold-out MSE for this configuration (synthetic_test_script.py:124-127).  The joint test log-probability
    # that train_exact_gp also records is NOT compared here: with noise 0.01 the 4000 x 4000 predictive covariance is numerically
    # singular and a CG solve at eval_cg_tolerance 5e-4 (the script's setting, as in GPyTorch) does not determine it -- the 1 % NLL
    # check lives in test_train_exact_gp_end_to_end_cfg1_shape (noise 0.05, eval tolerance 1e-4).
    assert np.isfinite(metrics["test_nll"])
    pred_var_ref = np.diag(ref_cov) + noise
    assert np.all(pred_var_ref > 0)


def test_experiment_runner_end_to_end(tmp_path):
    """the reference's UCI protocol through gp_experiment_runner.main -- spec file, fold split, train-set
    normalisation, solver flags (with --fast_pred: LOVE variances), train_exact_gp on the fused kernels, CSV on disk."""
    import gp_experiment_runner as runner
    X, y = synthetic(900, 5, seed=31, device="cpu")
    frame = runner.frame_from_array(np.concatenate([X.numpy(), y.numpy()[:, None]], axis=1))
    spec = tr.load_model_spec("additive_rp_J20_K1")
    spec["train_kwargs"].update(max_iter=12, check_conv=False)
    spec_path, out = tmp_path / "spec.json", tmp_path / "res.csv"
    spec_path.write_text(json.dumps(spec))
    torch.manual_seed(5)
    np.random.seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        table = runner.main(["-m", str(spec_path), "-d", "synthetic", "-o", str(out), "-s", "0.2", "--no_cv", "--fold", "2",
                             "--cg_tol", "0.01", "--eval_cg_tol", "0.001", "--fast_pred", "--skip_random_restart", "--device", "cuda:0",
                             "--error_repeats", "1"], datasets_override={"synthetic": frame})
    assert len(table) == 1 and "error" not in table.columns, table.get("error")
    row = table.iloc[0]
    assert row["fold"] == 2 and row["n"] == 900 and row["d"] == 5 and row["trained_epochs"] == 12
    assert np.isfinite(row["rmse"]) and row["rmse"] < 0.6 and np.isfinite(row["test_nll"]) and 0.3 < row["test_pred_frac_in_cr"] <= 1.0
    import pandas as pd
    saved = pd.read_csv(out)
    assert bool(saved["fast_pred_var"][0]) and saved["dataset"][0] == "synthetic" and abs(saved["cg_tol"][0] - 0.01) < 1e-12


@pytest.mark.parametrize("wrap", [False, True])
def test_additive_component_posteriors(wrap):
    """TestAdditivePredictions (test.py:384-457): per-component posteriors of an additive model -- equal for equal inputs,
    different otherwise, their means add up to the model's predictive mean -- then at a size where the solve runs through CG on
    the fused kernels, against the dense FP64 formula K_j* K^-1 y."""
    from gp_models import AdditiveExactGPModel, StrictlyAdditiveKernel
    kernel = StrictlyAdditiveKernel(2, RBFKernel)
    if wrap:
        kernel = ScaleKernel(kernel)
    lik = gpytorch.likelihoods.GaussianLikelihood()
    trainX, trainY = torch.tensor([[0., 0.]], device=DEV), torch.tensor([2.], device=DEV)
    model = AdditiveExactGPModel(trainX, trainY, lik, kernel).to(DEV)
    model.eval()
    equi = model.additive_pred(torch.tensor([[1., 1.]], device=DEV))
    diff = model.additive_pred(torch.tensor([[1., 0.]], device=DEV))
    assert abs(float(equi[0].mean[0]) - float(equi[1].mean[0])) < 1e-7
    assert abs(float(diff[0].mean[0]) - float(diff[1].mean[0])) > 1e-3
    with torch.no_grad():
        total = model(torch.tensor([[1., 0.]], device=DEV))
    combined = diff[0] + diff[1]
    assert abs(float(combined.mean[0]) - float(total.mean[0])) < 1e-6
    assert float(model.additive_pred(torch.tensor([[1., 0.], [1.1, 1.2]], device=DEV), group=1).variance.min()) > 0

    # n = 1500: CG on the symmetric tensor-core kernel, rectangular single-group products for the component means
    g = torch.Generator().manual_seed(3)
    X = torch.rand(1500, 2, generator=g).to(DEV) * 4 - 2
    y = (torch.sin(X[:, 0]) + torch.cos(2 * X[:, 1]) + 0.1 * torch.randn(1500, generator=g).to(DEV))
    Xt = torch.rand(64, 2, generator=g).to(DEV) * 4 - 2
    kernel = StrictlyAdditiveKernel(2, RBFKernel)
    kernel.initialize([1.0, 1.0], [0.7, 0.7])
    if wrap:
        kernel = ScaleKernel(kernel)
        kernel.outputscale = 1.7
    model = AdditiveExactGPModel(X, y, gpytorch.likelihoods.GaussianLikelihood(), kernel).to(DEV)
    model.eval()
    with settings.eval_cg_tolerance(1e-6), settings.max_cg_iterations(4000), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        preds = model.additive_pred(Xt)
        with torch.no_grad():
            total = model(Xt).mean
        var0 = preds[0].variance.cpu().numpy()
    assert len(preds) == 2
    summed = (preds[0] + preds[1]).mean
    assert rel(summed.cpu().numpy(), total.cpu().numpy()) < 1e-4
    # dense FP64 reference of component 0
    ls = 0.7
    s = 1.7 if wrap else 1.0
    Xd, Xtd, yd = X.cpu().double().numpy(), Xt.cpu().double().numpy(), y.cpu().double().numpy()
    k = lambda a, b: 0.5 * s * np.exp(-0.5 * ((a[:, None] - b[None, :]) / ls) ** 2)
    Kfull = k(Xd[:, 0], Xd[:, 0]) + k(Xd[:, 1], Xd[:, 1]) + float(model.likelihood.noise.detach()) * np.eye(1500)
    alpha = np.linalg.solve(Kfull, yd)
    ref_mean0 = k(Xtd[:, 0], Xd[:, 0]) @ alpha
    assert rel(preds[0].mean.cpu().numpy(), ref_mean0) < 2e-3, rel(preds[0].mean.cpu().numpy(), ref_mean0)
    c0 = k(Xtd[:, 0], Xd[:, 0])
    ref_var0 = np.diag(k(Xtd[:, 0], Xtd[:, 0]) - c0 @ np.linalg.solve(Kfull, c0.T))
    assert rel(var0, ref_var0) < 5e-3, rel(var0, ref_var0)
    assert tuple(preds[1].sample(torch.Size([3])).shape) == (3, 64)


def test_rp_model_equals_its_additive_conversion():
    """test.py:359-380: after one optimiser step the additive model over the projected inputs predicts what the RP model predicts"""
    from gp_models import convert_rp_model_to_additive_model
    g = torch.Generator().manual_seed(8)
    X, y = torch.randn(300, 2, generator=g).to(DEV), torch.randn(300, generator=g).to(DEV)
    Xt = torch.randn(40, 2, generator=g).to(DEV)
    Ws, bs = [torch.eye(2, 2) for _ in range(3)], [torch.zeros(2) for _ in range(3)]
    kernel = PolynomialProjectionKernel(3, 2, 2, RBFKernel, Ws, bs, learn_proj=False, weighted=True)
    lik = gpytorch.likelihoods.GaussianLikelihood()
    model = ExactGPModel(X, y, lik, kernel).to(DEV)
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    opt.zero_grad()
    loss = -mll(model(X), y)
    loss.backward()
    opt.step()
    add_model, projection = convert_rp_model_to_additive_model(model, True)
    add_model.eval()
    model.eval()
    with torch.no_grad():
        got = add_model(projection(Xt)).mean
        want = model(Xt).mean
    assert rel(got.cpu().numpy(), want.cpu().numpy()) < 1e-5


def test_synthetic_benchmark_point():
    """one point of a synthetic_test_script.py learning curve (BASELINE configs[0] at reduced size): the additive target in 6
    dimensions, 640 training points, the GAM model, trained with Adam through CG on the fused kernels; hold-out RMSE well below 1
    (the targets are standardised)"""
    import synthetic_test_script as sts
    torch.manual_seed(4)
    np.random.seed(4)
    ho_x = torch.rand(1000, 6) * 4 - 2
    ho_y = sts.additive(ho_x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mses, models, mlls = sts.benchmark_on_n_pts(640, sts.create_gam_model, sts.additive, ho_x, ho_y, repeats=1, max_iter=30,
                                                    return_model=True, device="cuda:0")
    assert len(mses) == 1 and len(models) == 1 and np.isfinite(mses[0])
    assert np.sqrt(mses[0]) < 0.3, mses
