"""GPU parity of the symmetric tensor-core forward (rpgp_mvm_sym_f32: tcgen05, 3xTF32 split, every kernel value used for
both out[i] and out[i']) against the FP64 oracle and against the SIMT forward kernel; tolerance 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import rpgp_oracle as orc
from rpgp import _lib

from parity_util import rel, sampled_oracle_check

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def setup(n, J, t, seed, spread=1.0, K=1):
    rng = np.random.RandomState(seed)
    Z = (rng.randn(n, J * K) * spread / np.sqrt(K)).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    V = rng.randn(n, t).astype(np.float32)
    lay = _lib.plan_layout(J, K)
    zp = _lib.pack_coords(torch.from_numpy(Z).to(DEV), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
    return Z, c, V, lay, zp, nlc


@pytest.mark.parametrize("n,J,t", [(50, 20, 11), (128, 20, 11), (129, 20, 3), (300, 20, 11), (777, 20, 1), (1000, 20, 16),
                                   (1024, 26, 16), (2500, 20, 11), (640, 7, 5), (900, 32, 11), (385, 3, 2)])
def test_sym_matches_oracle(n, J, t):
    Z, c, V, lay, zp, nlc = setup(n, J, t, seed=n + J)
    assert _lib.mvm_sym_supported(lay, t)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    ref = orc.kmv(Z, Z, c, J, 1, V)
    assert np.isfinite(got).all()
    assert rel(got, ref) < 1e-5, rel(got, ref)


@pytest.mark.parametrize("n,J,K,t", [(300, 20, 5, 11), (1000, 1, 20, 11), (777, 3, 2, 4), (1025, 10, 3, 16), (640, 2, 16, 1),
                                     (900, 40, 1, 11), (515, 90, 1, 3), (1300, 7, 8, 11), (260, 1, 32, 2)])
def test_sym_group_and_chunked_layouts_match_oracle(n, J, K, t):
    """K > 1 (one exponential per group of K coordinates) and J*K > 32 (several coordinate chunks, summed by the FP64 accumulators)."""
    Z, c, V, lay, zp, nlc = setup(n, J, t, seed=n + J + K, K=K)
    assert _lib.mvm_sym_supported(lay, t)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    ref = orc.kmv(Z, Z, c, J, K, V)
    assert np.isfinite(got).all()
    assert rel(got, ref) < 1e-5, rel(got, ref)


def test_sym_agrees_with_simt_forward_and_wide_spread():
    Z, c, V, lay, zp, nlc = setup(3000, 20, 11, seed=1, spread=9.5)
    Vd = torch.from_numpy(V).to(DEV)
    a = _lib.mvm_sym(zp, lay, nlc, Vd).cpu().numpy()
    b = _lib.mvm_fwd(zp, zp, lay, nlc, Vd).cpu().numpy()
    assert rel(a, b) < 2e-6, rel(a, b)
    assert rel(a, orc.kmv(Z, Z, c, 20, 1, V)) < 1e-5


def test_sym_block_ranges_sum_to_full():
    Z, c, V, lay, zp, nlc = setup(1500, 20, 11, seed=2)
    Vd = torch.from_numpy(V).to(DEV)
    full = _lib.mvm_sym(zp, lay, nlc, Vd).cpu().numpy().astype(np.float64)
    nb = (1500 + 127) // 128
    parts = sum(_lib.mvm_sym(zp, lay, nlc, Vd, block_range=(b0, b1)).cpu().numpy().astype(np.float64)
                for b0, b1 in [(0, 4), (4, 9), (9, nb)])
    assert rel(parts, full) < 2e-6
    torch.cuda.synchronize()


def test_sym_unsupported_shapes_are_reported():
    assert _lib.mvm_sym_supported(_lib.plan_layout(20, 5), 11)
    assert _lib.mvm_sym_supported(_lib.plan_layout(90, 1), 16)
    assert not _lib.mvm_sym_supported(_lib.plan_layout(20, 1), 17)
    assert not _lib.mvm_sym_supported(_lib.plan_layout(20, 1), 0)


# ---- K > 1 with the squared distances on the tensor cores (csrc/sym_tcd.cu) ------------------------------------------------------
@pytest.mark.parametrize("n,J,K,t", [(300, 20, 5, 11), (1000, 1, 20, 11), (1025, 10, 4, 16), (640, 2, 16, 1), (1300, 7, 8, 11),
                                     (260, 1, 24, 2), (3000, 3, 6, 5), (129, 9, 5, 3), (4000, 20, 5, 11), (2000, 17, 7, 4),
                                     (50, 3, 5, 2), (128, 8, 6, 11), (33, 1, 4, 1)])
def test_tensor_core_distances_match_oracle(n, J, K, t):
    """4 <= K <= 24: U = |z|^2 + |z'|^2 - 2 z.z' as one augmented inner product on tcgen05 (3xTF32), several groups per chunk,
    several chunks, partial last blocks; same 1e-5 bound as every other forward path."""
    Z, c, V, lay, zp, nlc = setup(n, J, t, seed=7 * n + J + K, K=K)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    ref = orc.kmv(Z, Z, c, J, K, V)
    assert np.isfinite(got).all()
    assert rel(got, ref) < 1e-5, rel(got, ref)


def _two_clusters(n, J, K, R2, seed, offset=0.0):
    """every within-cluster pair is near (k ~ 1) while the scaled |z|^2 ~ R2 is large: the worst case for the cancellation"""
    rng = np.random.RandomState(seed)
    R = np.sqrt(R2 / 0.72134752 / K)
    sign = np.where(rng.rand(n, 1) < 0.5, -1.0, 1.0)
    Z = (sign * R + 0.3 * rng.randn(n, J * K) / np.sqrt(K) + offset).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    lay = _lib.plan_layout(J, K)
    zp = _lib.pack_coords(torch.from_numpy(Z).to(DEV), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
    return Z, c, lay, zp, nlc


@pytest.mark.parametrize("J,K,R2,offset", [(20, 5, 128.0, 0.0), (1, 20, 128.0, 0.0), (8, 6, 64.0, 0.0), (20, 5, 32.0, 25.0)])
def test_tensor_core_distances_worst_case_radius(J, K, R2, offset):
    """tight clusters far from the origin (below the gate, so the tensor-core path runs), with and without a common offset
    (removed by the centring of the operand images); the systematic part of the error is what K.1 shows"""
    n, t = 3000, 11
    Z, c, lay, zp, nlc = _two_clusters(n, J, K, R2, seed=int(R2) + J, offset=offset)
    V = np.random.RandomState(5).randn(n, t).astype(np.float32)
    V[:, 0] = 1.0
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    ref = orc.kmv(Z, Z, c, J, K, V)
    assert rel(got, ref) < 1e-5, rel(got, ref)
    assert rel(got[:, 0], ref[:, 0]) < 1e-5, rel(got[:, 0], ref[:, 0])


def test_tensor_core_distances_gate_hands_over_to_direct_differences():
    """beyond the radius bound the pre-pass flag sends the launch to the direct-difference kernel: the result keeps that kernel's
    accuracy (the tensor-core distances would be ~4e-5 off here, tools/tcd_check.py adv)"""
    n, t, J, K = 3000, 11, 20, 5
    Z, c, lay, zp, nlc = _two_clusters(n, J, K, 2048.0, seed=11)
    V = np.random.RandomState(6).randn(n, t).astype(np.float32)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    assert rel(got, orc.kmv(Z, Z, c, J, K, V)) < 3e-6


def test_tensor_core_distances_block_ranges_sum_to_full():
    Z, c, V, lay, zp, nlc = setup(1500, 20, 11, seed=3, K=5)
    Vd = torch.from_numpy(V).to(DEV)
    full = _lib.mvm_sym(zp, lay, nlc, Vd).cpu().numpy().astype(np.float64)
    nb = (1500 + 127) // 128
    parts = sum(_lib.mvm_sym(zp, lay, nlc, Vd, block_range=(b0, b1)).cpu().numpy().astype(np.float64)
                for b0, b1 in [(0, 4), (4, 9), (9, nb)])
    assert rel(parts, full) < 2e-6
    assert rel(full, orc.kmv(Z, Z, c, 20, 5, V)) < 1e-5


# ---- at the benchmarked scale: sampled rows against the FP64 C oracle ---------------------------------------------
def _large(n, J, K, t, seed, spread):
    rng = np.random.RandomState(seed)
    Z = (rng.randn(n, J * K) * spread / np.sqrt(K)).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    V = rng.randn(n, t).astype(np.float32)
    lay = _lib.plan_layout(J, K)
    zp = _lib.pack_coords(torch.from_numpy(Z).to(DEV), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
    return c, V, lay, zp, nlc


@pytest.mark.parametrize("n,J,K,t,spread", [(200_000, 20, 1, 11, 2.0), (262_144 + 77, 26, 1, 16, 1.0), (200_000, 20, 5, 11, 2.0),
                                            (200_000, 1, 20, 11, 2.0)])
def test_sym_large_n_sampled_rows_match_f64_oracle(n, J, K, t, spread):
    """n >= 200k: ~1600 row blocks, the row side folds > 1500 tile epochs into Kahan-compensated FP32, the column side adds them with
    FP64 RED; both the direct-difference kernel (K = 1) and the distance-on-tensor-core kernel (K = 5, 20).  Norm-wise AND row-wise
    1e-5 on 128 sampled rows (first / last / random runs) against oracle_kmv_f64 on the identical packed coordinates."""
    c, V, lay, zp, nlc = _large(n, J, K, t, seed=n % 1000 + J + K, spread=spread)
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV))
    (norm_rel, row_rel, _), text = sampled_oracle_check(zp, lay, c, J, K, V, got)
    assert norm_rel < 1e-5 and row_rel < 1e-5, text


def test_sym_eight_uneven_rank_shares_sum_to_full():
    """single-GPU emulation of the 8-rank split of the unique block pairs (uneven shares, one of them empty): the partial products
    summed in FP64 equal the full product -- what the NCCL all-reduce of bench.py / rpgp.ops.kmv_partitioned computes"""
    n, J, t = 20_000 + 55, 20, 11
    c, V, lay, zp, nlc = _large(n, J, 1, t, seed=8, spread=1.5)
    Vd = torch.from_numpy(V).to(DEV)
    full = _lib.mvm_sym(zp, lay, nlc, Vd).double()
    nb = (n + 127) // 128
    cuts = [0, 3, 3, 40, 41, 77, 100, 140, nb]
    parts = sum(_lib.mvm_sym(zp, lay, nlc, Vd, block_range=(cuts[r], cuts[r + 1])).double() for r in range(8))
    e = rel(parts.cpu().numpy(), full.cpu().numpy())
    assert e < 2e-6, repr(e)
    (norm_rel, row_rel, _), text = sampled_oracle_check(zp, lay, c, J, 1, V, parts.float())
    assert norm_rel < 1e-5 and row_rel < 1e-5, text


@pytest.mark.parametrize("J,K", [(20, 5), (1, 20)])
def test_tensor_core_distances_at_the_gate_boundary(J, K):
    """tight clusters whose centred, scaled squared group norm sits just under the device-side gate (rms r2 <= 200): the
    distance-on-tensor-core kernel runs (checked from the same statistics the pre-pass computes) and must still hold 1e-5"""
    n, t = 3000, 11
    Z, c, lay, zp, nlc = _two_clusters(n, J, K, 185.0, seed=21 + K)
    plan = _lib.mvm_sym_distance_plan(lay)
    assert plan is not None
    zc = zp - zp.mean(dim=1, keepdim=True)
    r2 = torch.stack([(zc[ch, :, g * lay.KP:g * lay.KP + K] ** 2).sum(-1) for ch in range(lay.nchunks) for g in range(lay.G)
                      if ch * lay.G + g < J])
    rms = float((r2.double() ** 2).mean().sqrt())
    assert 150.0 < rms <= plan["bound"] and float(r2.max()) <= 10 * plan["bound"], (rms, float(r2.max()))   # inside the gate, near its edge
    V = np.random.RandomState(5).randn(n, t).astype(np.float32)
    V[:, 0] = 1.0
    got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
    ref = orc.kmv(Z, Z, c, J, K, V)
    e = rel(got, ref)
    assert e < 1e-5, repr(e)
    e0 = rel(got[:, :1], ref[:, :1])
    assert e0 < 1e-5, repr(e0)
