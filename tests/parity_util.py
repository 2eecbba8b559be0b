"""Helpers of the at-scale parity tests: sampled rows of a GPU product against the FP64 C oracle (oracle/kmv_oracle.c) applied to
EXACTLY the packed FP32 coordinates the kernels read.  Test infrastructure only."""
import numpy as np
import torch

from oracle import c_oracle
from rpgp import _lib


def sample_rows(n, count=128, seed=1):
    """first / last 32 rows (the last 128-row block is partial unless 128 | n) plus runs of 8 consecutive rows at random positions"""
    rng = np.random.RandomState(seed)
    rows = set(range(min(32, n))) | set(range(max(0, n - 32), n))
    while len(rows) < min(count, n):
        r0 = int(rng.randint(0, max(1, n - 8)))
        rows.update(range(r0, min(n, r0 + 8)))
    return np.array(sorted(rows)[:count], dtype=np.int64)


def natural_f64(zp, lay, J, K):
    """packed, pre-scaled planes (nchunks, n, CP) -> natural (n, J*K) float64 holding exactly the values the kernels see"""
    z = zp.cpu().numpy()
    nch, n, _ = z.shape
    g = z[:, :, :lay.G * lay.KP].reshape(nch, n, lay.G, lay.KP)[..., :K]
    g = np.transpose(g, (1, 0, 2, 3)).reshape(n, nch * lay.G, K)[:, :J, :]
    return np.ascontiguousarray(g.reshape(n, J * K), dtype=np.float64) / _lib.coord_scale()


def errors(got, ref):
    """(norm-wise relative error, largest row-wise relative error, largest absolute element error)"""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    diff = got - ref
    row = np.linalg.norm(diff.reshape(len(diff), -1), axis=1) / np.maximum(np.linalg.norm(ref.reshape(len(ref), -1), axis=1), 1e-300)
    return float(np.linalg.norm(diff) / max(np.linalg.norm(ref), 1e-300)), float(row.max()), float(np.abs(diff).max())


def describe(got, ref):
    e = errors(got, ref)
    return "norm-wise %.3g, max row-wise %.3g, max |element| %.3g (reference max |element| %.3g)" % (e + (float(np.abs(ref).max()),))


def sampled_oracle_check(zp, lay, c, J, K, V, got, count=128, seed=1):
    """errors of `got` (n x t GPU product, torch) on sampled rows against oracle_kmv_f64"""
    n = zp.shape[1]
    rows = sample_rows(n, count, seed)
    Zn = natural_f64(zp, lay, J, K)
    ref = c_oracle.kmv(Zn[rows], Zn, np.asarray(c, np.float64), J, K, np.asarray(V, np.float64), dtype=np.float64)
    g = got[torch.as_tensor(rows, device=got.device)].double().cpu().numpy()
    return errors(g, ref), describe(g, ref)


class Rel(float):
    """norm-wise relative error ||a - b|| / ||b|| as a float whose repr -- what pytest prints when `assert rel(a, b) < tol` fails --
    also carries the row-wise and element-wise figures"""

    def __new__(cls, a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        self = super().__new__(cls, np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        try:
            self.detail = describe(a.reshape(len(a), -1), b.reshape(len(b), -1)) if (a.ndim >= 1 and a.size) else ""
        except Exception:
            self.detail = ""
        return self

    def __repr__(self):
        return "%.3g [%s]" % (float(self), self.detail)

    __str__ = __repr__


def rel(a, b):
    return Rel(a, b)
