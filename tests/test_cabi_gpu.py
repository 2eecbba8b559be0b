"""GPU parity of the C ABI (librpgp.so) against the CPU oracle: forward K.V, quadratic-form gradients, rows, FP64.

Tolerances (BASELINE.json north_star): K.V <= 1e-5 relative in FP32 (norm-wise vs the FP64 oracle on identical FP32
inputs), <= 1e-10 for the FP64 build at small n; gradients <= 1e-4 relative.
"""
import numpy as np
import pytest
import torch

from oracle import rpgp_oracle as orc
from rpgp import _lib

pytestmark = pytest.mark.gpu


from parity_util import rel  # noqa: E402  (norm-wise error; its repr adds the row-wise / element-wise figures)


def make_problem(m, n, J, K, t, seed, spread=1.0, equal_c=False):
    rng = np.random.RandomState(seed)
    Z1 = (rng.randn(m, J * K) * spread).astype(np.float32)
    Z2 = (rng.randn(n, J * K) * spread).astype(np.float32)
    c = np.full(J, 0.7 / J, np.float32) if equal_c else (rng.rand(J).astype(np.float32) + 0.1)
    V = rng.randn(n, t).astype(np.float32)
    return Z1, Z2, c, V


def cuda_kmv(Z1, Z2, c, J, K, V, row_range=None):
    lay = _lib.plan_layout(J, K)
    dev = torch.device("cuda:0")
    z1p = _lib.pack_coords(torch.from_numpy(Z1).to(dev), lay)
    z2p = _lib.pack_coords(torch.from_numpy(Z2).to(dev), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(dev), lay)
    out = _lib.mvm_fwd(z1p, z2p, lay, nlc, torch.from_numpy(V).to(dev), row_range=row_range)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("m,n,J,K,t", [
    (2, 2, 3, 1, 1), (1, 1, 1, 1, 1), (257, 130, 20, 1, 11), (300, 777, 20, 1, 1), (512, 2000, 20, 1, 11),
    (100, 1000, 26, 1, 16), (64, 65, 7, 1, 5), (1000, 999, 33, 1, 8), (200, 300, 90, 1, 11), (130, 140, 20, 1, 32),
    (90, 300, 20, 1, 40), (150, 333, 1, 20, 11), (150, 333, 20, 5, 11), (70, 200, 3, 2, 3), (70, 200, 4, 3, 16),
    (50, 100, 2, 32, 2), (60, 120, 5, 8, 20),
])
def test_mvm_fwd_matches_oracle(m, n, J, K, t):
    Z1, Z2, c, V = make_problem(m, n, J, K, t, seed=m + n + J)
    ref = orc.kmv(Z1, Z2, c, J, K, V)
    got = cuda_kmv(Z1, Z2, c, J, K, V)
    assert got.shape == ref.shape
    assert rel(got, ref) < 1e-5, rel(got, ref)


def test_mvm_fwd_wide_spread_many_underflows():
    # Gaussian W with d=90 gives z ~ N(0, 90): most pairs underflow; must stay finite and accurate
    Z1, Z2, c, V = make_problem(300, 4000, 20, 1, 11, seed=5, spread=9.5)
    ref = orc.kmv(Z1, Z2, c, 20, 1, V)
    got = cuda_kmv(Z1, Z2, c, 20, 1, V)
    assert np.isfinite(got).all()
    assert rel(got, ref) < 1e-5, rel(got, ref)


def test_mvm_fwd_row_block_equals_full():
    Z1, Z2, c, V = make_problem(1000, 1000, 20, 1, 11, seed=9)
    full = cuda_kmv(Z1, Z2, c, 20, 1, V)
    blk = cuda_kmv(Z1, Z2, c, 20, 1, V, row_range=(300, 811))
    np.testing.assert_array_equal(blk, full[300:811])


def test_mvm_fwd_bit_reproducible_and_linear():
    Z1, Z2, c, V = make_problem(700, 5000, 20, 1, 11, seed=3)
    a = cuda_kmv(Z1, Z2, c, 20, 1, V)
    b = cuda_kmv(Z1, Z2, c, 20, 1, V)
    np.testing.assert_array_equal(a, b)
    V2 = np.random.RandomState(1).randn(*V.shape).astype(np.float32)
    s = cuda_kmv(Z1, Z2, c, 20, 1, (V + 2 * V2).astype(np.float32))
    assert rel(s, a.astype(np.float64) + 2 * cuda_kmv(Z1, Z2, c, 20, 1, V2)) < 1e-5


def test_mvm_large_n_accumulation():
    # 200k columns: the compensated two-level accumulation must hold 1e-5
    Z1, Z2, c, V = make_problem(256, 200_000, 20, 1, 11, seed=11)
    ref = orc.kmv(Z1, Z2, c, 20, 1, V, row_chunk=64)
    got = cuda_kmv(Z1, Z2, c, 20, 1, V)
    assert rel(got, ref) < 1e-5, rel(got, ref)


@pytest.mark.parametrize("m,n,J,K,t,sym", [
    (3, 2, 3, 1, 1, False), (257, 300, 20, 1, 11, False), (300, 300, 20, 1, 11, True), (128, 500, 26, 1, 16, False),
    (200, 200, 33, 1, 4, True), (150, 170, 1, 20, 11, False), (160, 160, 20, 5, 11, True), (90, 110, 3, 2, 20, False),
])
def test_quad_bwd_matches_oracle(m, n, J, K, t, sym):
    rng = np.random.RandomState(m * 3 + n)
    Z1, Z2, c, _ = make_problem(m, n, J, K, t, seed=m + J)
    if sym:
        Z2 = Z1
    L = rng.randn(m, t).astype(np.float32)
    R = rng.randn(n, t).astype(np.float32)
    dZ1_ref, dZ2_ref, dc_ref = orc.quad_form_grads(Z1, Z2, c, J, K, L, R)
    lay = _lib.plan_layout(J, K)
    dev = torch.device("cuda:0")
    z1p = _lib.pack_coords(torch.from_numpy(Z1).to(dev), lay)
    z2p = z1p if sym else _lib.pack_coords(torch.from_numpy(Z2).to(dev), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(dev), lay)
    Lt, Rt = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    dzp, g = _lib.quad_bwd(z1p, z2p, lay, nlc, Lt, Rt, symmetric=sym)
    torch.cuda.synchronize()
    # unpack: packed coordinate = scale * natural
    scale = _lib.coord_scale()
    dz = unpack(dzp.cpu().numpy(), lay) * scale
    want = dZ1_ref + dZ2_ref if sym else dZ1_ref
    assert rel(dz, want) < 1e-4, rel(dz, want)
    g = g.cpu().numpy()[:J]
    assert rel(g, dc_ref * c) < 1e-4, rel(g, dc_ref * c)


def unpack(zp, lay):
    nch, n, CP = zp.shape
    out = np.zeros((n, lay.J * lay.K), dtype=zp.dtype)
    for ch in range(nch):
        for g in range(lay.G):
            jg = ch * lay.G + g
            if jg >= lay.J:
                continue
            out[:, jg * lay.K:(jg + 1) * lay.K] = zp[ch, :, g * lay.KP:g * lay.KP + lay.K]
    return out


def test_pack_roundtrip_and_project():
    rng = np.random.RandomState(0)
    n, d, J, K = 1000, 13, 7, 3
    X = rng.randn(n, d).astype(np.float32)
    W = rng.randn(J * K, d).astype(np.float32)
    ell = (rng.rand(d) + 0.5).astype(np.float32)
    lay = _lib.plan_layout(J, K)
    dev = torch.device("cuda:0")
    Z = orc.scaled_projection(X, W, ell, prescale=True)
    zp = _lib.project(torch.from_numpy(X).to(dev), torch.from_numpy(W).to(dev), torch.from_numpy(1.0 / ell).to(dev),
                      None, lay).cpu().numpy()
    got = unpack(zp, lay) / _lib.coord_scale()
    ref = orc.scaled_projection(X, W, (1.0 / (1.0 / ell).astype(np.float32)), prescale=True)
    assert rel(got, ref) < 3e-7
    zp2 = _lib.pack_coords(torch.from_numpy(Z.astype(np.float32)).to(dev), lay).cpu().numpy()
    assert rel(unpack(zp2, lay) / _lib.coord_scale(), Z) < 3e-7
    # postscale
    ell2 = (rng.rand(J * K) + 0.5).astype(np.float32)
    zp3 = _lib.project(torch.from_numpy(X).to(dev), torch.from_numpy(W).to(dev), None,
                       torch.from_numpy(1.0 / ell2).to(dev), lay).cpu().numpy()
    ref3 = orc.scaled_projection(X, W, 1.0 / (1.0 / ell2).astype(np.float32), prescale=False)
    assert rel(unpack(zp3, lay) / _lib.coord_scale(), ref3) < 3e-7


def test_kernel_rows_and_f64_path():
    Z1, Z2, c, V = make_problem(40, 300, 6, 2, 7, seed=2)
    dev = torch.device("cuda:0")
    ref = orc.additive_rbf_dense(Z1, Z2, c, 6, 2)
    got = _lib.kernel_rows(torch.from_numpy(Z1).to(dev), torch.from_numpy(Z2).to(dev), torch.from_numpy(c).to(dev), 6, 2)
    assert rel(got.cpu().numpy(), ref) < 1e-6
    Z1d, Z2d, Vd, cd = (torch.from_numpy(a.astype(np.float64)).to(dev) for a in (Z1, Z2, V, c))
    got64 = _lib.kernel_rows(Z1d, Z2d, cd, 6, 2)
    assert rel(got64.cpu().numpy(), ref) < 1e-13
    kv = _lib.mvm_fwd_f64(Z1d, Z2d, cd, 6, 2, Vd)
    assert rel(kv.cpu().numpy(), orc.kmv(Z1, Z2, c, 6, 2, V)) < 1e-10
    L = np.random.RandomState(4).randn(40, 7)
    dZ1, g = _lib.quad_bwd_f64(Z1d, Z2d, cd, 6, 2, torch.from_numpy(L).to(dev), Vd)
    dZ1_ref, _, dc_ref = orc.quad_form_grads(Z1, Z2, c, 6, 2, L, V)
    assert rel(dZ1.cpu().numpy(), dZ1_ref) < 1e-10
    assert rel(g.cpu().numpy(), dc_ref * c) < 1e-10


def test_kmv_host_entry_point():
    rng = np.random.RandomState(7)
    n, d, J, t = 1500, 10, 20, 11
    X = (rng.rand(n, d) * 4 - 2).astype(np.float32)
    W = (rng.randn(J, d)).astype(np.float32)
    ell = np.ones(d, np.float32)
    c = np.full(J, np.log(2.0) / J, np.float32)
    V = rng.randn(n, t).astype(np.float32)
    got = _lib.kmv_host(X, None, W, J, 1, 1.0 / ell, None, c, V, diag_add=0.5)
    Z = orc.scaled_projection(X, W, ell, prescale=True)
    ref = orc.kmv(Z, Z, c, J, 1, V, diag_add=0.5)
    assert rel(got, ref) < 1e-5, rel(got, ref)
    # rectangular (prediction-shaped) call
    X2 = (rng.rand(700, d) * 4 - 2).astype(np.float32)
    got2 = _lib.kmv_host(X2, X, W, J, 1, 1.0 / ell, None, c, V)
    ref2 = orc.kmv(orc.scaled_projection(X2, W, ell, True), Z, c, J, 1, V)
    assert rel(got2, ref2) < 1e-5, rel(got2, ref2)


def test_error_reporting():
    lay = _lib.plan_layout(20, 1)
    dev = torch.device("cuda:0")
    z = torch.zeros((1, 10, lay.CP), device=dev)
    nlc = torch.zeros((lay.CP,), device=dev)
    with pytest.raises(RuntimeError):
        _lib.mvm_fwd(z.cpu(), z, lay, nlc, torch.zeros((10, 3), device=dev))
    with pytest.raises(RuntimeError, match="status"):
        bad = _lib.Layout(20, 1, 7, 1, 1, 7)
        _lib.pack_log2c(torch.ones(20, device=dev), bad)


# ---- the host-buffer path as a persistent plan (rpgp_plan_*; bench.py's e2e) ------------------------------------------------------
@pytest.mark.parametrize("n,d,J,K,t", [(3000, 10, 20, 1, 11), (700, 6, 5, 1, 3), (2600, 12, 4, 5, 11), (1500, 8, 26, 1, 16), (1200, 7, 5, 1, 20),
                                       (1030, 5, 40, 1, 2)])
def test_host_plan_matches_oracle_and_block_shares_sum_to_full(n, d, J, K, t):
    """set_operator (H2D + projection) then kmv (H2D of V, symmetric product, + sigma^2 V, D2H) through the plan handle: the product
    against the FP64 oracle, the one-shot rpgp_kmv_host_f32, and three uneven rank shares summed on the host"""
    rng = np.random.RandomState(n + J)
    X = rng.randn(n, d).astype(np.float32)
    W = (rng.randn(J * K, d) / np.sqrt(K * d) * 2.0).astype(np.float32)
    ell = (0.5 + rng.rand(d)).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    V = rng.randn(n, t).astype(np.float32)
    Z = orc.scaled_projection(X, W, ell, prescale=True)
    ref = orc.kmv(Z, Z, c, J, K, V) + 0.25 * V.astype(np.float64)
    plan = _lib.HostPlan(n, d, J, K, t, device=0)
    try:
        plan.set_operator(X, W, 1.0 / ell, None, c)
        got = plan.kmv(V, diag_add=0.25)
        assert rel(got, ref) < 1e-5, rel(got, ref)
        again = plan.kmv(V[:, :max(1, t // 2)], diag_add=0.25)               # fewer right-hand sides through the same plan
        assert rel(again, ref[:, :max(1, t // 2)]) < 1e-5
        one_shot = _lib.kmv_host(X, None, W, J, K, 1.0 / ell, None, c, V, diag_add=0.25, device=0)
        assert rel(got, one_shot) < 2e-6, rel(got, one_shot)
        nb = (n + 127) // 128
        cuts = [0, nb // 3, nb // 3 + 1, nb]
        parts = sum(plan.kmv(V, block_range=(cuts[r], cuts[r + 1])).astype(np.float64) for r in range(3))
        assert rel(parts + 0.25 * V, ref) < 1e-5
        rows = plan.kmv(V, diag_add=0.25, row_range=(100, 357))
        assert rows.shape == (257, t) and rel(rows, ref[100:357]) < 1e-5
        plan.set_operator(X, W, 2.0 / ell, None, c)                            # new hyper-parameters through the same buffers
        Z2 = orc.scaled_projection(X, W, ell / 2.0, prescale=True)
        assert rel(plan.kmv(V), orc.kmv(Z2, Z2, c, J, K, V)) < 1e-5
    finally:
        plan.close()
