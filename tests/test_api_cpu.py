"""CPU: the host-side mirror of the reference's kernel / model API (structure, parameter names, init semantics, specs),
the C-ABI library's exported symbols, and the loud failure of the product path without CUDA.

Pins G7 of SURVEY.md §8c: parameter counts / frozen flags (test.py:126-134, 172-175, 188-190), init semantics
(test.py:302-319), projection-cache semantics (test.py:156-165, 177-183), lengthscale of frozen base kernels
(test.py:597-599).
"""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import training_routines as tr
from gp_models.kernels import (GeneralizedProjectionKernel, MemoryEfficientGamKernel, PolynomialProjectionKernel,
                               RPPolyKernel, ScaledProjectionKernel, StrictlyAdditiveKernel)
from gp_models.models import ExactGPModel
from rpgp import _lib
from rpgp import gp as gpytorch
from rpgp.gp.kernels import AdditiveStructureKernel, LazyEvaluatedKernelTensor, RBFKernel, ScaleKernel
from rpgp.lazy import RPAdditiveLazyTensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ----------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rpgp.h")).read()
    declared = set(re.findall(r"\b(rpgp_[a-z0-9_]+)\s*\(", header))
    declared -= {"rpgp_layout", "rpgp_status"}
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "librpgp.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().rpgp_version() >= 100


def test_plan_handle_argument_checks_need_no_gpu():
    """rpgp_plan_create validates its shape arguments before touching the device; a NULL plan is rejected by every entry point"""
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.rpgp_plan_create(0, 3, 2, 1, 4, 0, None, ctypes.byref(h))
    assert rc == 1 and b"bad shape" in lib.rpgp_last_error() and not h
    assert lib.rpgp_plan_kmv_begin(None, None, 1, 0, 0) == 1
    assert lib.rpgp_plan_kmv_end(None, 0.0, None, 0, 0) == 1
    assert lib.rpgp_plan_set_operator(None, None, None, None, None, None) == 1
    assert lib.rpgp_plan_device_out(None) is None
    assert lib.rpgp_plan_destroy(None) == 0


def test_layout_planner_host_logic():
    for (J, K) in [(20, 1), (26, 1), (1, 1), (90, 1), (33, 1), (1, 20), (20, 5), (3, 2), (4, 3), (2, 32), (5, 8)]:
        lay = _lib.plan_layout(J, K)
        assert lay.J == J and lay.K == K and lay.CP % 4 == 0 and 4 <= lay.CP <= 32
        assert lay.nchunks * lay.G >= J and lay.KP >= K and lay.G * lay.KP <= lay.CP
        assert (lay.KP == 1) == (K == 1)
    assert _lib.plan_layout(20, 1).key() == (20, 1, 20, 1, 1, 20)
    assert _lib.plan_layout(90, 1).nchunks == 3
    lay = _lib.plan_layout(20, 1)
    assert [_lib.padded_rhs(lay, t) for t in (1, 4, 5, 11, 12, 16, 17, 32, 33)] == [4, 4, 8, 12, 12, 16, 32, 32, 0]
    assert _lib.padded_rhs(lay, 11, backward=True) == 12 and _lib.max_rhs(lay, True) == 16
    with pytest.raises(RuntimeError, match="exceeds"):
        _lib.plan_layout(1, 40)
    assert abs(_lib.coord_scale() - np.sqrt(0.5 * np.log2(np.e))) < 1e-15


def test_product_path_fails_loudly_without_cuda():
    x = torch.randn(10, 3)
    kernel = ScaleKernel(RBFKernel())
    op = kernel(x, x).evaluate_kernel()
    assert isinstance(op, RPAdditiveLazyTensor)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op.evaluate()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op.matmul(torch.randn(10, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op._matmul(torch.randn(10, 2))


# ---- kernel structure ----------------------------------------------------------------------------------------------------
def make_poly(J=3, k=2, d=6, **kw):
    Ws = [torch.eye(d, k) for _ in range(J)]
    bs = [torch.zeros(k) for _ in range(J)]
    return PolynomialProjectionKernel(J, k, d, RBFKernel, Ws, bs, activation=None, **kw)


def test_poly_projection_kernel_parameters_and_flags():
    J, k = 3, 2
    kernel = make_poly(J, k, learn_proj=False, weighted=False)
    params = dict(kernel.named_parameters())
    assert len(params) == J * (k + 1) + 2                                   # lengthscales + mixins + W + b
    assert params["kernel.kernels.0.base_kernel.kernels.0.raw_lengthscale"] == 0
    for name in ["kernel.kernels.0.base_kernel.kernels.1.raw_lengthscale", "kernel.kernels.2.base_kernel.kernels.0.raw_lengthscale",
                 "kernel.kernels.2.raw_outputscale", "kernel.kernels.0.raw_outputscale"]:
        assert name in params
    assert not kernel.kernel.kernels[0].raw_outputscale.requires_grad
    assert not kernel.projection_module.weight.requires_grad and not kernel.projection_module.bias.requires_grad
    assert abs(kernel.kernel.kernels[0].outputscale.item() - 1 / J) < 1e-6
    weighted = make_poly(J, k, weighted=True)
    assert weighted.kernel.kernels[0].raw_outputscale.requires_grad
    learn = make_poly(J, k, learn_proj=True)
    assert learn.projection_module.weight.requires_grad and learn.projection_module.bias.requires_grad
    with pytest.raises(ValueError):
        PolynomialProjectionKernel(J, k, 6, RBFKernel, [torch.eye(6, k)] * J, [torch.zeros(k)] * J, activation="relu")
    with pytest.raises(NotImplementedError):
        make_poly(ski=True)


def test_poly_projection_kernel_lowering_and_cache():
    torch.manual_seed(0)
    x = torch.randn(12, 6)
    kernel = make_poly(3, 2)
    assert kernel.last_x1 is None and kernel.cached_projections is None
    out = kernel(x, x)
    assert isinstance(out, LazyEvaluatedKernelTensor)
    op = out.evaluate_kernel()
    assert isinstance(op, RPAdditiveLazyTensor) and op.symmetric
    assert (op.J, op.K) == (3, 2) and op.shape == (12, 12)
    np.testing.assert_allclose(op.c.detach().numpy(), [1 / 3] * 3, rtol=1e-6)
    ls = float(torch.nn.functional.softplus(torch.zeros(())))
    np.testing.assert_allclose(op.Z1.detach().numpy(), np.tile(x[:, :2].numpy() / ls, (1, 3)), rtol=1e-6)
    assert kernel.last_x1 is not None and kernel.cached_projections is not None
    kernel(x[:10], x[:10]).evaluate_kernel()
    assert kernel.last_x1.numel() == x[:10].numel()
    learn = make_poly(3, 2, learn_proj=True)
    learn(x, x).evaluate_kernel()
    assert learn.last_x1 is None and learn.cached_projections is None
    # rectangular: x2 is projected separately
    rect = kernel(x[:5], x).evaluate_kernel()
    assert not rect.symmetric and rect.shape == (5, 12)
    # diag of the symmetric operator is sum_j c_j
    np.testing.assert_allclose(kernel(x, x, diag=True).detach().numpy(), np.ones(12), rtol=1e-6)


def test_generalized_kernel_initialize_semantics():
    torch.manual_seed(1)
    proj = torch.nn.Linear(6, 5, bias=False)
    kernel = GeneralizedProjectionKernel([2, 3], 6, RBFKernel, proj, weighted=True)
    kernel.initialize([1.0, 1.0], [0.1, 0.1])
    np.testing.assert_allclose([k.outputscale.item() for k in kernel.kernel.kernels], [0.5, 0.5], rtol=1e-5)
    for bk in kernel.base_kernels:
        np.testing.assert_allclose(bk.lengthscale.item(), 0.1, rtol=1e-5)
    assert len(kernel.base_kernels) == 5 and len(kernel.scale_kernels) == 2
    add = kernel.to_additive_kernel()
    assert add.groups == [[0, 1], [2, 3, 4]] and add.kernel is kernel.kernel
    sa = StrictlyAdditiveKernel(4, RBFKernel)
    op = sa(torch.randn(7, 4)).evaluate_kernel()
    assert (op.J, op.K) == (4, 1)


def test_scaled_projection_kernel_structure_and_lowering():
    x = torch.tensor([[1., 2., 3.], [1.1, 2.2, 3.3]])
    for prescale in (True, False):
        kbase = RBFKernel()
        kbase.initialize(lengthscale=torch.tensor([1.]))
        base = AdditiveStructureKernel(kbase, 3)
        proj = torch.nn.Linear(3, 3, bias=False)
        proj.weight.data = torch.eye(3)
        k = ScaledProjectionKernel(proj, base, prescale=prescale, ard_num_dims=3)
        k.initialize(lengthscale=torch.tensor([1., 2., 3.]))
        np.testing.assert_allclose(k.lengthscale.detach().numpy(), [[1., 2., 3.]], rtol=1e-6)
        assert not proj.weight.requires_grad and not kbase.raw_lengthscale.requires_grad and k.raw_lengthscale.requires_grad
        np.testing.assert_allclose(k.base_kernel.base_kernel.lengthscale.detach().numpy(), [[1.]], rtol=1e-6)
        op = k(x, x).evaluate_kernel()
        assert (op.J, op.K) == (3, 1) and op.symmetric
        np.testing.assert_allclose(op.Z1.detach().numpy(), x.numpy() / np.array([[1., 2., 3.]]), rtol=1e-6)
    k2 = ScaledProjectionKernel(torch.nn.Linear(3, 3, bias=False), base, prescale=True, ard_num_dims=3, learn_proj=True)
    assert k2.projection_module.weight.requires_grad
    gam = MemoryEfficientGamKernel()
    opg = gam(x, x).evaluate_kernel()
    assert (opg.J, opg.K) == (3, 1) and abs(gam.lengthscale.item() - np.log(2)) < 1e-6
    np.testing.assert_allclose(opg.c.numpy(), np.ones(3))


# ---- factories / specs ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec,J,K,csum", [
    ("additive_rp_J20_K1", 20, 1, 1.0), ("additive_rp_prescale_J20", 20, 1, 1.0), ("additive_rp_postscale_J20", 20, 1, 1.0),
    ("additive_spread_prescale_Jd", 7, 1, 7.0), ("additive_rp_prescale_J1_K20", 1, 20, 1.0),
    ("additive_rp_prescale_J20_K5", 20, 5, 1.0), ("additive_spread_prescale_J20", 20, 1, 1.0),
])
def test_model_specs_build_and_lower(spec, J, K, csum):
    torch.manual_seed(0)
    np.random.seed(0)
    s = tr.load_model_spec(spec)
    assert set(s) == {"kind", "model_kwargs", "train_kwargs"}
    X, y = torch.randn(30, 7), torch.randn(30)
    kw = {k: (7 if v == "d" else v) for k, v in s["model_kwargs"].items()}
    model, likelihood = tr.create_exact_gp(X, y, s["kind"], **kw)
    assert isinstance(model, ExactGPModel) and abs(likelihood.noise.item() - 1.0) < 1e-5
    assert abs(model.covar_module.outputscale.item() - np.log(2)) < 1e-6      # outer ScaleKernel, softplus(0)
    op = model.covar_module(X).evaluate_kernel()
    assert (op.J, op.K) == (J, K) and op.symmetric and op.shape == (30, 30)
    np.testing.assert_allclose(op.c.sum().item(), csum * np.log(2), rtol=1e-5)
    trainable = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    assert "likelihood.raw_noise" in trainable and "mean_module.constant" in trainable
    assert "covar_module.raw_outputscale" in trainable
    if s["kind"] == "additive_rp":
        assert "covar_module.base_kernel.raw_lengthscale" in trainable
        assert not any("projection_module" in n for n in trainable)
        want = 7 if kw.get("prescale") else J * K
        assert model.covar_module.base_kernel.raw_lengthscale.shape == (1, want)
    prior = dict((n, p) for n, p, _ in gpytorch.mlls.ExactMarginalLogLikelihood(likelihood, model).named_priors())
    assert list(prior) == ["likelihood.noise_prior"]


def test_spread_projections_are_orthonormal_and_unknown_kinds_raise():
    torch.manual_seed(0)
    np.random.seed(0)
    k = tr.create_additive_rp_kernel(7, 7, space_proj=True, prescale=True, batch_kernel=False, mem_efficient=True)
    W = k.projection_module.weight.detach().numpy()
    np.testing.assert_allclose(W @ W.T, np.eye(7), atol=1e-5)
    with pytest.raises(ValueError):
        tr.create_additive_rp_kernel(7, 3, k=2)                                # k > 1 needs batch_kernel=False
    X, y = torch.randn(5, 3), torch.randn(5)
    with pytest.raises(ValueError):
        tr.create_exact_gp(X, y, "nonsense", noise_prior=False)
    with pytest.raises(NotImplementedError):
        tr.create_exact_gp(X, y, "sgpr", noise_prior=False)
    with pytest.raises(NotImplementedError):
        tr.create_rp_poly_kernel(3, 1, 2, ski=True)


def test_priors_and_constraints():
    lp = gpytorch.priors.SmoothedBoxPrior(1e-4, 10, sigma=0.01)
    from oracle import rpgp_oracle as orc
    for v in (0.5, 1.0, 9.99, 10.05, 1e-5):
        np.testing.assert_allclose(lp.log_prob(torch.tensor(v, dtype=torch.float64)).item(), orc.smoothed_box_log_prob(v), rtol=1e-10)
    lik = gpytorch.likelihoods.GaussianLikelihood()
    lik.noise = torch.tensor([0.37])
    assert abs(lik.noise.item() - 0.37) < 1e-6
    assert abs(gpytorch.likelihoods.GaussianLikelihood().noise.item() - (np.log(2) + 1e-4)) < 1e-6
    m = gpytorch.means.ConstantMean()
    assert m(torch.randn(4, 3)).shape == (4,) and m.constant.item() == 0


def test_distance_on_tensor_core_plan_and_base_layouts():
    """host-side planning that needs no GPU: the chunking of the K > 1 tensor-core-distance path (csrc/sym_tcd.cu: six coordinates +
    two partial norms per k-step of eight, at most eight k-steps per chunk) and the layouts of the non-RBF base kernels"""
    from rpgp import _lib
    plan = _lib.mvm_sym_distance_plan(_lib.plan_layout(20, 5))
    assert plan == dict(groups_per_chunk=7, ksteps_per_group=1, lines=2, nchunks=3, bound=plan["bound"]) and plan["bound"] > 0
    plan = _lib.mvm_sym_distance_plan(_lib.plan_layout(1, 20))
    assert (plan["groups_per_chunk"], plan["ksteps_per_group"], plan["lines"], plan["nchunks"]) == (1, 4, 1, 1)
    plan = _lib.mvm_sym_distance_plan(_lib.plan_layout(8, 6))
    assert (plan["groups_per_chunk"], plan["ksteps_per_group"], plan["lines"], plan["nchunks"]) == (8, 1, 2, 1)
    plan = _lib.mvm_sym_distance_plan(_lib.plan_layout(5, 12))
    assert (plan["groups_per_chunk"], plan["ksteps_per_group"], plan["lines"], plan["nchunks"]) == (3, 2, 2, 2)
    for J, K in [(20, 1), (10, 3), (1, 30)]:          # K = 1, K below the cross-over, K too wide: direct differences only
        assert _lib.mvm_sym_distance_plan(_lib.plan_layout(J, K)) is None
    # non-RBF base kernels: the same layouts as the RBF (K = 1: one coordinate per projection), the symmetric tensor-core kernel but
    # not its distance-on-tensor-core variant, fewer right-hand-side widths
    rbf, mat = _lib.plan_layout(20, 1), _lib.plan_layout(20, 1, _lib.BASE_MATERN15)
    assert (rbf.base, rbf.KP, rbf.G, rbf.CP) == (0, 1, 20, 20)
    assert mat.base == 1 and (mat.KP, mat.G, mat.CP, mat.nchunks, mat.K) == (1, 20, 20, 1, 1)
    assert _lib.mvm_sym_supported(rbf, 11) and _lib.mvm_sym_supported(mat, 11) and not _lib.mvm_sym_supported(mat, 17)
    assert _lib.mvm_sym_distance_plan(_lib.plan_layout(20, 5, _lib.BASE_INVERSE_MQ)) is None
    assert _lib.padded_rhs(mat, 11, False) == 16 and _lib.padded_rhs(rbf, 11, False) == 12
    imq = _lib.plan_layout(3, 4, _lib.BASE_INVERSE_MQ)
    assert imq.base == 2 and imq.KP >= 4
    cos1, cos4 = _lib.plan_layout(20, 1, _lib.BASE_COSINE), _lib.plan_layout(3, 4, _lib.BASE_COSINE)
    assert cos1.base == 3 and (cos1.KP, cos1.CP) == (1, 20) and _lib.mvm_sym_supported(cos1, 11)
    assert not _lib.mvm_sym_supported(cos4, 11)       # the cosine kernel with K > 1 takes the rectangular kernel
    with pytest.raises(RuntimeError):
        _lib.plan_layout(3, 4, 7)


def test_random_projections_preserve_distances_like_the_reference_checks():
    """TestRPGenerator (test.py:52-109): mean relative distortion of pairwise distances below 10 % for a 100 -> 1000 dimensional
    projection drawn by gen_rp, for every distribution the reference tests; spherical columns have equal norms (:77)"""
    import rp
    torch.manual_seed(0)
    data = torch.randn(40, 100)
    ref = torch.cdist(data, data)
    for dist in ("gaussian", "sphere", "bernoulli", "uniform"):
        W = rp.gen_rp(100, 1000, dist=dist)
        assert tuple(W.shape) == (100, 1000)
        proj = data.matmul(W)
        err = (ref - torch.cdist(proj, proj)).abs().mean() / ref.abs().mean()
        assert float(err) < 0.1, (dist, float(err))
    W = rp.gen_rp(100, 1000, dist="sphere")
    assert abs(W[:, 0].norm().item() - W[:, 1].norm().item()) < 1e-5


def test_strictly_additive_and_grouped_additive_kernels():
    """TestStrictlyAdditiveKernel / TestAdditiveKernel (test.py:330-357): structure of the component kernels, `initialize`, and
    that the grouping of the features matters -- read off the operator the kernel lowers to (no evaluation needed)"""
    from gp_models.kernels import CustomAdditiveKernel
    d = 5
    kernel = StrictlyAdditiveKernel(d, RBFKernel)
    assert isinstance(kernel.kernel, gpytorch.kernels.AdditiveKernel) and len(kernel.kernel.kernels) == d
    assert isinstance(kernel.kernel.kernels[0].base_kernel, RBFKernel)
    kernel.initialize([0.1, 0.1], [0.1, 0.1])
    x = torch.randn(7, d)
    op = kernel.forward(x, x)
    assert isinstance(op, RPAdditiveLazyTensor) and (op.J, op.K) == (d, 1) and op.symmetric
    np.testing.assert_allclose(op.c.detach().numpy(), np.full(d, 1.0 / d), rtol=1e-6)        # mixins normalised to sum to one
    np.testing.assert_allclose(op.Z1.detach().numpy(), (x / 0.1).numpy(), rtol=1e-5)         # lengthscale 0.1 on every feature

    k1 = CustomAdditiveKernel([[1, 2], [0, 3]], 4, RBFKernel)
    k2 = CustomAdditiveKernel([[0, 1], [2, 3]], 4, RBFKernel)
    assert isinstance(k1.kernel.kernels[0].base_kernel.kernels[0], RBFKernel)
    xx = torch.tensor([[0., 1., 2., 3.], [0., 1., 4., 5.]])
    o1, o2 = k1.forward(xx, xx), k2.forward(xx, xx)
    assert (o1.J, o1.K) == (2, 2) and (o2.J, o2.K) == (2, 2)

    def dense(o):      # the operator's definition, in torch (CPU): sum_j c_j exp(-1/2 |z_i^(j) - z_i'^(j)|^2)
        z = o.Z1.detach().reshape(o.Z1.shape[0], o.J, o.K)
        sq = ((z[:, None] - z[None, :]) ** 2).sum(-1)
        return (o.c.detach() * torch.exp(-0.5 * sq)).sum(-1)

    assert abs(float(dense(o1)[0, 1]) - float(dense(o2)[0, 1])) > 1e-3                       # test.py:350-357


def test_additive_models_structure_and_conversion():
    """gp_models/models.py:23-86,110-125 on the host: kernel-type checks, groups, and the RP -> additive conversion shares the
    component kernels / likelihood / mean and carries the projected training inputs (numerics: tests/test_model_gpu.py)"""
    from gp_models import AdditiveExactGPModel, CustomAdditiveKernel, ProjectedAdditiveExactGPModel, convert_rp_model_to_additive_model
    x, y = torch.randn(6, 4), torch.randn(6)
    lik = gpytorch.likelihoods.GaussianLikelihood()
    groups = [[1, 2], [0, 3]]
    m = AdditiveExactGPModel(x, y, lik, CustomAdditiveKernel(groups, 4, RBFKernel))
    assert m.get_groups() == groups
    ms = AdditiveExactGPModel(x, y, lik, ScaleKernel(StrictlyAdditiveKernel(4, RBFKernel)))
    assert ms.get_groups() == [[0], [1], [2], [3]]
    for bad in (RBFKernel(), ScaleKernel(RBFKernel())):
        with pytest.raises(ValueError):
            AdditiveExactGPModel(x, y, lik, bad)
        with pytest.raises(ValueError):
            ProjectedAdditiveExactGPModel(x, y, lik, bad)
    Ws, bs = [torch.eye(4, 2) for _ in range(3)], [torch.zeros(2) for _ in range(3)]
    rp_kernel = PolynomialProjectionKernel(3, 2, 4, RBFKernel, Ws, bs, learn_proj=False, weighted=True)
    rp_model = ProjectedAdditiveExactGPModel(x, y, lik, rp_kernel)
    add_model, proj = rp_model.get_corresponding_additive_model()
    assert isinstance(add_model, AdditiveExactGPModel) and proj is rp_kernel.projection_module
    assert add_model.covar_module.kernel is rp_kernel.kernel and add_model.likelihood is lik and add_model.mean_module is rp_model.mean_module
    assert add_model.get_groups() == [[0, 1], [2, 3], [4, 5]] and tuple(add_model.train_inputs[0].shape) == (6, 6)
    np.testing.assert_allclose(add_model.train_inputs[0].numpy(), proj(x).detach().numpy())
    wrapped = ExactGPModel(x, y, lik, ScaleKernel(rp_kernel))
    wrapped.covar_module.outputscale = 2.5
    add2 = convert_rp_model_to_additive_model(wrapped, return_proj=False)
    assert isinstance(add2.covar_module, ScaleKernel) and abs(float(add2.covar_module.outputscale.detach()) - 2.5) < 1e-6


def test_multivariate_normal_sum_and_sampling():
    from rpgp.gp.distributions import MultivariateNormal
    a = MultivariateNormal(torch.tensor([1.0, 2.0]), torch.tensor([[2.0, 0.5], [0.5, 1.0]]))
    b = MultivariateNormal(torch.tensor([0.5, -1.0]), torch.eye(2))
    s = a + b
    np.testing.assert_allclose(s.mean.numpy(), [1.5, 1.0])
    np.testing.assert_allclose(s.covariance_matrix.numpy(), [[3.0, 0.5], [0.5, 2.0]])
    torch.manual_seed(0)
    draws = a.sample(torch.Size([20000]))
    assert tuple(draws.shape) == (20000, 2)
    np.testing.assert_allclose(draws.mean(0).numpy(), [1.0, 2.0], atol=0.05)
    np.testing.assert_allclose(np.cov(draws.numpy().T), [[2.0, 0.5], [0.5, 1.0]], atol=0.08)
    assert tuple(a.sample().shape) == (2,)
