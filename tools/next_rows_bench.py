"""Measurements for the 'next' rows of SURVEY §8(f) on one B200 (CUDA events, 3 warm-ups):
  f4  forward K.V with the Matern-1.5 / inverse-multiquadric / cosine base kernels against RBF on the same SIMT kernel (cfg2 shape)
  f3  predictive variances: LOVE root (Lanczos steps) + variances of n* test points, and one exact batch of test points
Prints one JSON line per measurement."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "randomly-projected-additive-gps_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from rpgp import _lib, lazy, ops  # noqa: E402
from rpgp.gp import settings  # noqa: E402
from rpgp.solver.lanczos import lanczos_root_inv  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def base_kernels(n=100_000, J=20, t=11):
    g = torch.Generator().manual_seed(0)
    Z = torch.randn(n, J, generator=g).to(DEV)
    c = torch.full((J,), 0.05, device=DEV)
    V = torch.randn(n, t, generator=g).to(DEV)
    for base, name in ((0, "rbf"), (1, "matern15"), (2, "inverse_mq"), (3, "cosine")):
        p = ops.Packed(Z, J, 1, base)
        nlc = ops.pack_weights(c, p.lay)
        ms = timed(lambda: _lib.mvm_fwd(p.zp, p.zp, p.lay, nlc, V))
        print(json.dumps({"row": "f4", "what": "forward K.V, SIMT kernel, base=%s" % name, "n": n, "J": J, "t": t, "KP": p.lay.KP,
                          "ms": ms, "pair_evals_per_s": n * n * J / (ms * 1e-3)}), flush=True)
        if _lib.mvm_sym_supported(p.lay, t):
            ms = timed(lambda: _lib.mvm_sym(p.zp, p.lay, nlc, V))
            print(json.dumps({"row": "f4", "what": "symmetric tcgen05 kernel, base=%s" % name, "n": n, "J": J, "t": t, "KP": p.lay.KP,
                              "G": p.lay.G, "nchunks": p.lay.nchunks, "ms": ms, "pair_evals_per_s": n * n * J / (ms * 1e-3)}), flush=True)


def predictive(n=50_000, nt=4096, J=20, rank=20):
    g = torch.Generator().manual_seed(1)
    Z = torch.randn(n, J, generator=g).to(DEV)
    Zt = torch.randn(nt, J, generator=g).to(DEV)
    c = torch.full((J,), 1.0 / J, device=DEV)
    train = lazy.AddedDiagLazyTensor(lazy.RPAdditiveLazyTensor(Z, None, c, J, 1), torch.tensor(0.5, device=DEV))
    cross = lazy.RPAdditiveLazyTensor(Zt, Z, c, J, 1)
    tt = lazy.RPAdditiveLazyTensor(Zt, None, c, J, 1)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        init = cross._transpose_nonbatch()._matmul(torch.full((nt, 1), 1.0 / nt, device=DEV))
        root = lanczos_root_inv(train._matmul, init, rank)
        torch.cuda.synchronize()
        t_root = time.perf_counter() - t0
        cov = lazy.PredictiveCovarLazyTensor(tt, cross, train, root=root)
        t0 = time.perf_counter()
        var = cov.diag()
        torch.cuda.synchronize()
        t_var = time.perf_counter() - t0
        print(json.dumps({"row": "f3", "what": "LOVE: Lanczos root of K^-1 + variances of all test points", "n": n, "n_test": nt,
                          "rank": int(root.shape[1]), "root_s": t_root, "variances_s": t_var, "min_var": float(var.min())}), flush=True)
        exact = lazy.PredictiveCovarLazyTensor(tt, cross, train)
        with settings.eval_cg_tolerance(0.01), settings.variance_batch_size(64), settings.max_preconditioner_size(0):
            sub = lazy.PredictiveCovarLazyTensor(lazy.RPAdditiveLazyTensor(Zt[:64].contiguous(), None, c, J, 1),
                                                 lazy.RPAdditiveLazyTensor(Zt[:64].contiguous(), Z, c, J, 1), train)
            t0 = time.perf_counter()
            v64 = sub.diag()
            torch.cuda.synchronize()
            t_exact = time.perf_counter() - t0
        print(json.dumps({"row": "f3", "what": "exact variances, one batch of 64 test points (multi-RHS CG, eval_cg_tolerance 0.01)",
                          "n": n, "seconds": t_exact, "love_vs_exact_max_abs_diff": float((var[:64] - v64).abs().max())}), flush=True)
        del exact


if __name__ == "__main__":
    import sys
    for part in ((base_kernels,) if "base" in sys.argv[1:] else (base_kernels, predictive)):
        try:
            part()
        except Exception as e:  # keep whatever was measured
            print(json.dumps({"error": "%s: %s" % (part.__name__, e)}), flush=True)
