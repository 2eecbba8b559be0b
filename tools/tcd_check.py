"""Accuracy and timing of the distance-on-tensor-core symmetric kernel (csrc/sym_tcd.cu) against the FP64 oracle and against
the direct-difference kernel (RPGP_SYM_TCD=0, separate process).  Usage: python tools/tcd_check.py [acc|time] ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "randomly-projected-additive-gps_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import rpgp_oracle as orc  # noqa: E402
from rpgp import _lib  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def setup(n, J, K, t, seed, spread=1.0):
    rng = np.random.RandomState(seed)
    Z = (rng.randn(n, J * K) * spread / np.sqrt(K)).astype(np.float32)
    c = (rng.rand(J) + 0.1).astype(np.float32)
    V = rng.randn(n, t).astype(np.float32)
    lay = _lib.plan_layout(J, K)
    zp = _lib.pack_coords(torch.from_numpy(Z).to(DEV), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
    return Z, c, V, lay, zp, nlc


def acc():
    cases = [(300, 20, 5, 11, 1.0), (1000, 1, 20, 11, 1.0), (1025, 10, 4, 16, 1.0), (640, 2, 16, 1, 1.0), (1300, 7, 8, 11, 1.0),
             (260, 1, 30, 2, 1.0), (2000, 20, 5, 11, 0.3), (2000, 20, 5, 11, 2.0), (2000, 20, 5, 11, 3.0), (2000, 1, 20, 11, 2.0),
             (3000, 3, 6, 5, 1.0), (129, 9, 5, 3, 1.0), (4000, 20, 5, 11, 1.0), (4000, 20, 5, 11, 3.0), (4000, 20, 5, 11, 5.0),
             (4000, 20, 5, 11, 9.5), (4000, 1, 20, 11, 5.0), (4000, 1, 20, 11, 9.5), (4000, 1, 20, 11, 20.0), (4000, 8, 6, 11, 6.0),
             (8000, 20, 5, 1, 9.5), (8000, 20, 5, 1, 4.0)]
    if len(sys.argv) > 2:
        cases = cases[-int(sys.argv[2]):]
    for n, J, K, t, spread in cases:
        Z, c, V, lay, zp, nlc = setup(n, J, K, t, seed=n + J + K, spread=spread)
        got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
        ref = orc.kmv(Z, Z, c, J, K, V)
        simt = _lib.mvm_fwd(zp, zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
        zs = Z.reshape(n, J, K).astype(np.float64) * 0.84932180028801904
        maxsq = (zs ** 2).sum(-1).max()
        # share of the result that comes from off-diagonal entries (what the cancellation error can touch)
        diag_only = (c.sum() * V)
        print(f"n={n} J={J} K={K} t={t} spread={spread}: sym rel {rel(got, ref):.2e}  simt rel {rel(simt, ref):.2e}  max|z|^2 {maxsq:.1f}  "
              f"offdiag share {rel(ref, diag_only):.2f} finite {np.isfinite(got).all()}", flush=True)


def adversarial():
    """two tight clusters at +-R per group: every within-cluster pair is near (k ~ 1) while |z|^2 ~ R^2 is large -- the worst case
    for the cancellation in |z|^2 + |z'|^2 - 2 z.z'.  Run with RPGP_SYM_TCD_AMAX=1e9 to see the raw tensor-core error."""
    for J, K in [(20, 5), (1, 20), (8, 6)]:
        for R2 in [8.0, 24.0, 64.0, 128.0, 256.0, 512.0, 2048.0]:
            n, t = 3000, 11
            rng = np.random.RandomState(int(R2) + J)
            # natural coordinates: scaled |z|^2 = 0.7213 * |z|^2
            R = np.sqrt(R2 / 0.72134752 / K)
            sign = np.where(rng.rand(n, 1) < 0.5, -1.0, 1.0)
            Z = (sign * R + 0.3 * rng.randn(n, J * K) / np.sqrt(K)).astype(np.float32)
            c = (rng.rand(J) + 0.1).astype(np.float32)
            V = rng.randn(n, t).astype(np.float32)
            lay = _lib.plan_layout(J, K)
            zp = _lib.pack_coords(torch.from_numpy(Z).to(DEV), lay)
            nlc = _lib.pack_log2c(torch.from_numpy(c).to(DEV), lay)
            got = _lib.mvm_sym(zp, lay, nlc, torch.from_numpy(V).to(DEV)).cpu().numpy()
            ones = _lib.mvm_sym(zp, lay, nlc, torch.ones(n, 1, device=DEV)).cpu().numpy()
            ref = orc.kmv(Z, Z, c, J, K, V)
            ref1 = orc.kmv(Z, Z, c, J, K, np.ones((n, 1), np.float32))
            print(f"adversarial J={J} K={K} R2={R2}: K.V rel {rel(got, ref):.2e}   K.1 rel {rel(ones, ref1):.2e} (systematic part)", flush=True)


def timing(n, J, K, t=11, reps=3):
    Z, c, V, lay, zp, nlc = setup(n, J, K, t, seed=1)
    Vd = torch.from_numpy(V).to(DEV)
    for _ in range(2):
        _lib.mvm_sym(zp, lay, nlc, Vd)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        _lib.mvm_sym(zp, lay, nlc, Vd)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    pe = n * n * J / (ms * 1e-3)
    print(f"time n={n} J={J} K={K} t={t} TCD={os.environ.get('RPGP_SYM_TCD', '1')}: {ms:.2f} ms  {pe:.3e} pair-evals/s "
          f"(extrapolated n=1M: {ms * (1e6 / n) ** 2 / 1e3:.2f} s)", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "acc":
        acc()
    elif mode == "adv":
        adversarial()
    else:
        n, J, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        timing(n, J, K)
