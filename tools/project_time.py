"""Timing of the projection kernels (forward on tcgen05, vector-Jacobian product) against the HBM roof: algorithmic bytes = 4 n d (X) + the
output (packed planes and / or natural rows); MEASURED_PEAKS.json hbm_gbs is the denominator."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "randomly-projected-additive-gps_b200"))
from rpgp import _lib  # noqa: E402

dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6555.8


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (name, n, d, J, K) in [("cfg4", 1_000_000, 90, 20, 1), ("cfg5b", 1_000_000, 90, 20, 5), ("cfg5a", 1_000_000, 90, 1, 20),
                           ("cfg3", 400_000, 26, 26, 1), ("cfg2", 100_000, 20, 20, 1)]:
    lay = _lib.plan_layout(J, K)
    X = torch.randn(n, d, device=dev)
    W = torch.randn(J * K, d, device=dev)
    ell = torch.ones(d, device=dev)
    ms = timeit(lambda: _lib.project(X, W, ell, None, lay))
    zp = _lib.project(X, W, ell, None, lay)
    byts = 4.0 * (n * d + zp.numel())
    print("%-6s project (packed)        n=%d d=%d J=%d K=%d: %.3f ms  %.0f GB/s = %.1f %% of the measured HBM roof (%.0f GB/s)"
          % (name, n, d, J, K, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / peak, peak))
    ms2 = timeit(lambda: _lib.project2(X, W, ell, None, lay, packed=False, natural=True))
    byts2 = 4.0 * (n * d + n * J * K)
    print("%-6s project (natural)       : %.3f ms  %.0f GB/s = %.1f %%" % (name, ms2, byts2 / ms2 / 1e6, 100 * byts2 / ms2 / 1e6 / peak))
    dZ = torch.randn(n, J * K, device=dev)
    ms3 = timeit(lambda: _lib.project_bwd(X, dZ))
    print("%-6s project_bwd (dZ^T X)    : %.3f ms  %.0f GB/s = %.1f %%" % (name, ms3, byts2 / ms3 / 1e6, 100 * byts2 / ms3 / 1e6 / peak))
    Xs = X  # cuBLAS reference for the same product (what the reference's nn.Linear runs)
    ms4 = timeit(lambda: torch.nn.functional.linear(Xs / ell, W))
    print("%-6s torch (x / l) @ W^T     : %.3f ms" % (name, ms4))
    del X, W, zp, dZ
