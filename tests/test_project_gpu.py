"""GPU parity of the projection Z^ = (X / l) W^T on the tensor cores (rpgp_project_f32 / rpgp_project2_f32: tcgen05 kind::tf32, split
precision) and of its vector-Jacobian product (rpgp_project_bwd_f32) against the FP64 oracle (oracle/rpgp_oracle.py scaled_projection,
restating gp_models/kernels/scaled_projection_kernel.py:21-37), and of the kernel classes that call them."""
import numpy as np
import pytest
import torch

from oracle import rpgp_oracle as orc
from rpgp import _lib, ops

from parity_util import rel

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def unpack(zp, lay):
    nch, n, _ = zp.shape
    g = zp[:, :, :lay.G * lay.KP].reshape(nch, n, lay.G, lay.KP)[..., :lay.K]
    return np.transpose(g, (1, 0, 2, 3)).reshape(n, nch * lay.G, lay.K)[:, :lay.J, :].reshape(n, lay.J * lay.K)


@pytest.mark.parametrize("n,d,J,K,mode", [(1000, 10, 20, 1, "pre"), (5000, 90, 20, 1, "pre"), (3001, 90, 20, 5, "pre"), (777, 26, 26, 1, "post"),
                                          (130, 128, 7, 16, "none"), (2000, 33, 3, 2, "both"), (128, 8, 1, 1, "pre"), (129, 1, 2, 1, "post"),
                                          (40000, 20, 20, 1, "pre"), (300, 64, 1, 20, "pre")])
def test_projection_on_tensor_cores_matches_oracle(n, d, J, K, mode):
    """packed planes and natural rows of one launch against the FP64 projection of the same FP32 inputs: the accuracy of an FP32 GEMM
    (what the reference's nn.Linear is): norm-wise 1e-6 -- the tensor core adds the exact products of the split parts into an FP32
    accumulator that truncates, 2 d / 8 additions deep, measured 5e-7 at d = 90 -- and every element within 4e-6 of sum |x||w|;
    padding positions exactly zero; partial last tile.  (RPGP_PROJECT_TC=0 selects the FP64-accumulating SIMT kernel: 6e-8.)"""
    rng = np.random.RandomState(n + d + J)
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(J * K, d) / np.sqrt(K)).astype(np.float32)
    pre = (0.5 + rng.rand(d)).astype(np.float32) if mode in ("pre", "both") else None
    post = (0.5 + rng.rand(J * K)).astype(np.float32) if mode in ("post", "both") else None
    lay = _lib.plan_layout(J, K)
    assert _lib.project_tc_supported(d, lay)
    t = lambda a: None if a is None else torch.from_numpy(a).to(DEV)      # noqa: E731
    zp, zn = _lib.project2(t(X), t(W), t(pre), t(post), lay)
    ref = (X.astype(np.float64) * (1.0 if pre is None else pre.astype(np.float64))) @ W.astype(np.float64).T
    if post is not None:
        ref = ref * post.astype(np.float64)
    got_n = zn.cpu().numpy()
    got_p = unpack(zp.cpu().numpy(), lay) / _lib.coord_scale()
    assert rel(got_n, ref) < 1e-6, rel(got_n, ref)
    assert rel(got_p, ref) < 1e-6, rel(got_p, ref)
    scale = np.abs(X).astype(np.float64) @ np.abs(W.astype(np.float64) * (1.0 if pre is None else pre)).T
    if post is not None:
        scale = scale * post
    assert np.all(np.abs(got_n - ref) <= 4e-6 * scale + 1e-30)
    # padding positions of the packed planes are exact zeros (the kernels rely on it)
    zpn = zp.cpu().numpy()
    mask = np.zeros((lay.nchunks, lay.CP), bool)
    for ch in range(lay.nchunks):
        for g in range(lay.G):
            if ch * lay.G + g < J:
                mask[ch, g * lay.KP:g * lay.KP + K] = True
    assert np.all(zpn[~np.broadcast_to(mask[:, None, :], zpn.shape)] == 0.0)
    # the single-output entry point gives the same planes
    zp1 = _lib.project(t(X), t(W), t(pre), t(post), lay)
    assert torch.equal(zp1, zp)


def test_projected_coordinates_feed_the_product_within_tolerance():
    """K.V on coordinates projected by the tensor-core kernel against the oracle's K.V on FP64-projected coordinates: 1e-5"""
    rng = np.random.RandomState(3)
    n, d, J, t = 3000, 90, 20, 11
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(J, d) / np.sqrt(d) * 1.5).astype(np.float32)
    ell = (0.8 + 0.4 * rng.rand(d)).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    V = rng.randn(n, t).astype(np.float32)
    lay = _lib.plan_layout(J, 1)
    zp = _lib.project(torch.from_numpy(X).to(DEV), torch.from_numpy(W).to(DEV), torch.from_numpy(1.0 / ell).to(DEV), None, lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    Z = orc.scaled_projection(X, W, ell, prescale=True)
    assert rel(got, orc.kmv(Z, Z, c, J, 1, V)) < 1e-5


@pytest.mark.parametrize("n,d,JK", [(1000, 10, 20), (5003, 90, 100), (257, 128, 112), (31, 3, 1), (40000, 26, 26), (2000, 7, 200)])
def test_projection_vjp_matches_oracle(n, d, JK):
    rng = np.random.RandomState(n + JK)
    X = rng.randn(n, d).astype(np.float32)
    dZ = rng.randn(n, JK).astype(np.float32)
    got = _lib.project_bwd(torch.from_numpy(X).to(DEV), torch.from_numpy(dZ).to(DEV)).cpu().numpy()
    ref = dZ.astype(np.float64).T @ X.astype(np.float64)
    assert rel(got, ref) < 2e-6, rel(got, ref)
    again = _lib.project_bwd(torch.from_numpy(X).to(DEV), torch.from_numpy(dZ).to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(got, again)          # fixed-order reduction: bit-reproducible


@pytest.mark.parametrize("mode", ["pre", "post", "both"])
def test_projection_autograd_matches_float64_reference(mode):
    """gradients of a scalar function of Z with respect to W, pre_inv and post_inv through rpgp.ops.project against torch float64
    autograd of ((x * pre) W^T) * post: 1e-4 (BASELINE.json gradient tolerance)"""
    torch.manual_seed(0)
    n, d, JK = 2000, 24, 30
    x = torch.randn(n, d, device=DEV)
    W = torch.randn(JK, d, device=DEV, requires_grad=True)
    pre = (0.5 + torch.rand(d, device=DEV)).requires_grad_(mode in ("pre", "both"))
    post = (0.5 + torch.rand(JK, device=DEV)).requires_grad_(mode in ("post", "both"))
    T = torch.randn(n, JK, device=DEV)
    Z = ops.project(x, W, pre if mode in ("pre", "both") else None, post if mode in ("post", "both") else None)
    ((Z * T).sum() + (Z ** 2).sum() * 0.01).backward()
    xd, Wd, Td = x.double(), W.detach().double().requires_grad_(True), T.double()
    pd = pre.detach().double().requires_grad_(True)
    qd = post.detach().double().requires_grad_(True)
    Zd = xd * pd if mode in ("pre", "both") else xd
    Zd = Zd @ Wd.t()
    Zd = Zd * qd if mode in ("post", "both") else Zd
    ((Zd * Td).sum() + (Zd ** 2).sum() * 0.01).backward()
    assert rel(Z.detach().cpu().numpy(), Zd.detach().cpu().numpy()) < 1e-6
    assert rel(W.grad.cpu().numpy(), Wd.grad.cpu().numpy()) < 1e-4
    if mode in ("pre", "both"):
        assert rel(pre.grad.cpu().numpy(), pd.grad.cpu().numpy()) < 1e-4
    if mode in ("post", "both"):
        assert rel(post.grad.cpu().numpy(), qd.grad.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("prescale", [True, False])
def test_scaled_projection_kernel_uses_the_library_projection(prescale):
    """ScaledProjectionKernel._scaled_projection (scaled_projection_kernel.py:21-27) runs the library's projection kernel -- no
    cuBLAS GEMM -- forward and backward, and agrees with nn.Linear + div in float64"""
    from gp_models.kernels import ScaledProjectionKernel
    from rpgp.gp.kernels import RBFKernel, ScaleKernel
    torch.manual_seed(1)
    n, d, JK = 1500, 12, 8
    x = torch.randn(n, d, device=DEV)
    pm = torch.nn.Linear(d, JK, bias=False).to(DEV)
    kern = ScaledProjectionKernel(pm, ScaleKernel(RBFKernel()), prescale=prescale, ard_num_dims=d if prescale else JK, learn_proj=True).to(DEV)
    kern.lengthscale = torch.linspace(0.5, 1.5, d if prescale else JK, device=DEV).reshape(1, -1)
    with _lib.timing() as tm:
        z = kern._scaled_projection(x)
        z.square().sum().backward()
        calls = tm.totals()
    assert calls.get("project", (0, 0))[0] == 1 and calls.get("project_bwd", (0, 0))[0] == 1, calls
    ls = kern.lengthscale.detach().double()
    Wd = pm.weight.detach().double().requires_grad_(True)
    raw = kern.raw_lengthscale.detach().double().requires_grad_(True)
    lsd = torch.nn.functional.softplus(raw)
    zd = (x.double() / lsd) @ Wd.t() if prescale else (x.double() @ Wd.t()) / lsd
    zd.square().sum().backward()
    assert rel(z.detach().cpu().numpy(), zd.detach().cpu().numpy()) < 1e-6
    assert rel(pm.weight.grad.cpu().numpy(), Wd.grad.cpu().numpy()) < 1e-4
    assert rel(kern.raw_lengthscale.grad.cpu().numpy(), raw.grad.cpu().numpy()) < 1e-4
    assert float((ls - lsd.detach()).abs().max()) < 1e-6
