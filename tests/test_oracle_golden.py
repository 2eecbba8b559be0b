"""CPU: pin the oracle (and the host-side rp.py) against the reference's golden vectors.

G1-G5 are the reference's own known-answer tests (test.py:533-573, 625-635, 640-680; numbers restated in SURVEY.md
§8c); tests/golden/*.npz were produced by running the reference's rp.py and GAMFunction (tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import rpgp_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_g1_g2_manual_rescale_known_answer():
    # test.py:533-573: x=[[1,2,3],[1.1,2.2,3.3]], W=I3, ell=[1,2,3], base RBF ls=1, J=3  =>  K = 3*k_RBF(x[:,0])
    x = np.array([[1., 2., 3.], [1.1, 2.2, 3.3]])
    ell = np.array([1., 2., 3.])
    want = np.array([[3.0, 2.985037326813], [2.985037326813, 3.0]])
    for prescale in (True, False):
        Z = orc.scaled_projection(x, np.eye(3), ell, prescale)
        K = orc.additive_rbf_dense(Z, Z, 1.0, 3, 1)
        np.testing.assert_allclose(K, want, rtol=1e-7)  # the reference asserts in FP32 (assert_allclose default rtol)
        np.testing.assert_allclose(K, 3 * np.exp(-0.5 * (x[:, :1] - x[:, :1].T) ** 2), rtol=1e-13)


def test_g3_memory_efficient_gam_default_lengthscale():
    # test.py:625-635: default raw_lengthscale 0 => ls = softplus(0) = ln 2
    x = np.array([[1., 2., 3.], [1.1, 2.2, 3.3]])
    ls = orc.softplus(0.0)
    assert abs(float(ls) - 0.6931471805599453) < 1e-15
    K = orc.gam_forward(x, x, np.full(1, ls))
    np.testing.assert_allclose(K, [[3.0, 2.859465122223], [2.859465122223, 3.0]], atol=1e-6)  # test.py:635 atol
    Z = x / ls
    np.testing.assert_allclose(orc.additive_rbf_dense(Z, Z, 1.0, 3, 1), K, rtol=1e-13)


def test_g4_g5_gamfunction_forward_backward():
    g = np.load(os.path.join(GOLD, "gam_g4_g5.npz"))
    ls = orc.softplus(g["raw_lengthscale"])
    np.testing.assert_allclose(ls, [1.313261687518, 3.048587351574, 2.126928011043], rtol=1e-12)
    np.testing.assert_allclose(ls, g["lengthscale"], rtol=1e-14)
    K = orc.gam_forward(g["x1"], g["x2"], ls)
    np.testing.assert_allclose(K, [[1.443414077343, 1.345462806615], [2.647681385118, 1.042931441187],
                                   [1.870386513267, 1.818452765552]], rtol=1e-12)
    np.testing.assert_allclose(K, g["K"], rtol=1e-14)
    dx1, dx2, dls = orc.gam_backward(g["x1"], g["x2"], ls, np.ones_like(K))
    np.testing.assert_allclose(dx1, g["dx1"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(dx2, g["dx2"], rtol=1e-12, atol=1e-15)
    # chain to the raw parameter: d softplus(r)/dr = sigmoid(r)
    draw = dls / (1.0 + np.exp(-g["raw_lengthscale"]))
    np.testing.assert_allclose(draw, g["draw"], rtol=1e-12)
    np.testing.assert_allclose(draw, [0.961318151498, 0.306416532810, 1.092260991964], rtol=1e-11)
    np.testing.assert_allclose(dx1[0], [0.128904983052, 0.010753976122, -0.008347863220], rtol=1e-9)
    np.testing.assert_allclose(dx2[1], [0.124008534327, 0.078986264915, -0.376758554137], rtol=1e-9)


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_gam_random_cases_match_reference(idx):
    g = np.load(os.path.join(GOLD, "gam_random.npz"))
    for tag, tol in (("f64", 1e-12), ("f32", 2e-5)):
        pre = "c%d_%s_" % (idx, tag)
        x1, x2, ell, L, R = (g[pre + k].astype(np.float64) for k in ("x1", "x2", "ell", "L", "R"))
        K = orc.gam_forward(x1, x2, ell)
        np.testing.assert_allclose(K, g[pre + "K"], rtol=tol, atol=tol)
        dx1, dx2, dell = orc.gam_backward(x1, x2, ell, L @ R.T)
        scale = max(1.0, np.abs(g[pre + "dx1"]).max())
        np.testing.assert_allclose(dx1, g[pre + "dx1"], rtol=tol * 50, atol=tol * 50 * scale)
        np.testing.assert_allclose(dx2, g[pre + "dx2"], rtol=tol * 50, atol=tol * 50 * scale)
        np.testing.assert_allclose(dell, g[pre + "dell"], rtol=tol * 200, atol=tol * 200 * np.abs(g[pre + "dell"]).max())


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_canonical_operator_equals_gamfunction(idx):
    """the (Z, c, J, K) normal form with Z = x/ell, c = 1, J = d, K = 1 IS the reference's GAMFunction,
    and the quadratic-form gradient formulas reproduce its backward."""
    g = np.load(os.path.join(GOLD, "gam_random.npz"))
    pre = "c%d_f64_" % idx
    x1, x2, ell, L, R = (g[pre + k] for k in ("x1", "x2", "ell", "L", "R"))
    d = x1.shape[1]
    ell_d = np.broadcast_to(ell, (d,))
    Z1, Z2 = x1 / ell_d, x2 / ell_d
    np.testing.assert_allclose(orc.additive_rbf_dense(Z1, Z2, 1.0, d, 1), g[pre + "K"], rtol=1e-12)
    dZ1, dZ2, dc = orc.quad_form_grads(Z1, Z2, 1.0, d, 1, L, R)
    np.testing.assert_allclose(dZ1 / ell_d, g[pre + "dx1"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dZ2 / ell_d, g[pre + "dx2"], rtol=1e-9, atol=1e-11)
    # d/d ell_k = sum_i dZ1[i,k] * (-x1[i,k]/ell_k^2) + same for x2
    dell = -(dZ1 * x1).sum(0) / ell_d ** 2 - (dZ2 * x2).sum(0) / ell_d ** 2
    if ell.size == 1:
        dell = dell.sum(keepdims=True)
    np.testing.assert_allclose(dell, g[pre + "dell"], rtol=1e-9, atol=1e-10)
    # K.V from the dense matrix
    V = R
    np.testing.assert_allclose(orc.kmv(Z1, Z2, 1.0, d, 1, V, row_chunk=7), g[pre + "K"] @ V, rtol=1e-12)


def test_multi_dim_groups_product_structure():
    # G6 restated on synthetic x (test.py:136-154): J=3 groups of k=2 identity coordinates == (1/3)*3*k(x0)k(x1)
    rng = np.random.RandomState(0)
    x = rng.randn(30, 5)
    ls = float(orc.softplus(0.0))
    Z = np.concatenate([x[:, :2] / ls] * 3, axis=1)
    K = orc.additive_rbf_dense(Z, Z, 1.0 / 3, 3, 2)
    k0 = np.exp(-0.5 * ((x[:, :1] - x[:, :1].T) / ls) ** 2)
    k1 = np.exp(-0.5 * ((x[:, 1:2] - x[:, 1:2].T) / ls) ** 2)
    np.testing.assert_allclose(K, k0 * k1, rtol=1e-12)


def test_rp_gen_rp_matches_reference_draws():
    import rp
    g = np.load(os.path.join(GOLD, "gen_rp.npz"))
    for dist in rp.RP_DISTRIBUTIONS:
        for (d, k) in [(10, 1), (20, 1), (90, 5), (7, 3)]:
            torch.manual_seed(1234)
            got = rp.gen_rp(d, k, dist).numpy()
            np.testing.assert_allclose(got, g["%s_%d_%d" % (dist, d, k)], rtol=1e-6, atol=1e-7)
    torch.manual_seed(7)
    W = torch.cat([rp.gen_rp(10, 1, "gaussian") for _ in range(20)], dim=1).t().numpy()
    np.testing.assert_array_equal(W, g["weight_J20_d10"])
    with pytest.raises(ValueError):
        rp.gen_rp(3, 1, "nope")


def test_rp_space_equally_matches_reference():
    import rp
    g = np.load(os.path.join(GOLD, "space_equally.npz"))
    np.random.seed(42)
    out, loss = rp.space_equally(torch.from_numpy(g["gs_in"]).clone(), lr=0.1, niter=5000)
    assert loss is None
    np.testing.assert_allclose(out.numpy(), g["gs_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out.numpy() @ out.numpy().T, np.eye(6), atol=1e-6)  # orthonormal when J <= d
    out2, loss2 = rp.space_equally(torch.from_numpy(g["gd_in"]).clone(), lr=0.1, niter=300)
    np.testing.assert_allclose(out2.numpy(), g["gd_out"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(np.linalg.norm(out2.numpy(), axis=1), 1.0, atol=1e-6)  # test.py:466-468
    np.testing.assert_allclose(loss2.numpy(), g["gd_loss"], rtol=1e-3)
    assert float(loss2) != 0.0  # test.py:484-491 (J > d cannot be orthogonal)


def test_exact_mll_gradients_against_finite_differences():
    rng = np.random.RandomState(3)
    n, J = 25, 4
    Z = rng.randn(n, J)
    c = rng.rand(J) + 0.2
    y = rng.randn(n)
    noise = 0.3

    def nmll(Zv, cv, nv, mu):
        Kh = orc.additive_rbf_dense(Zv, Zv, cv, J, 1) + nv * np.eye(n)
        return orc.exact_mll_dense(Kh, y, mu)[0] * n

    dZ, dc, dnoise, dmean = orc.exact_mll_grads_dense(Z, c, J, 1, noise, y, 0.1)
    eps = 1e-6
    Zp = Z.copy(); Zp[3, 2] += eps
    Zm = Z.copy(); Zm[3, 2] -= eps
    assert abs((nmll(Zp, c, noise, 0.1) - nmll(Zm, c, noise, 0.1)) / (2 * eps) - dZ[3, 2]) < 1e-6
    cp = c.copy(); cp[1] += eps
    cm = c.copy(); cm[1] -= eps
    assert abs((nmll(Z, cp, noise, 0.1) - nmll(Z, cm, noise, 0.1)) / (2 * eps) - dc[1]) < 1e-6
    assert abs((nmll(Z, c, noise + eps, 0.1) - nmll(Z, c, noise - eps, 0.1)) / (2 * eps) - dnoise) < 1e-6
    assert abs((nmll(Z, c, noise, 0.1 + eps) - nmll(Z, c, noise, 0.1 - eps)) / (2 * eps) - dmean) < 1e-6


def test_inverse_multiquadric_matches_reference_postprocess_function():
    """fixtures from the reference's own `postprocess_inverse_mq` (imq_kernel.py:8-9) applied the way InverseMQKernel.forward
    does (:17-22): pins base kernel 2 of the oracle (and through it of the CUDA kernels, tests/test_base_kernels_gpu.py)"""
    g = np.load(os.path.join(GOLD, "imq.npz"))
    np.testing.assert_allclose(orc.base_f(2, g["grid_sq"]), g["grid_k"], rtol=1e-14)
    for idx in range(3):
        x1, x2, ls = g["c%d_x1" % idx], g["c%d_x2" % idx], g["c%d_ls" % idx]
        d = x1.shape[1]
        K = orc.additive_rbf_dense(x1 / ls, x2 / ls, [1.0], 1, d, base=2)       # one group of d coordinates
        np.testing.assert_allclose(K, g["c%d_K" % idx], rtol=1e-12)
