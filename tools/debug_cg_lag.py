"""Debugging aid: linear_cg with and without the lagged convergence check on the same system."""
import os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "randomly-projected-additive-gps_b200"))
from rpgp.solver import linear_cg
from rpgp.gp import settings
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
n, t = 900, 200
X = torch.rand(n, 5, device=dev, generator=g) * 4 - 2
d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
K = 1.0 / torch.sqrt(1.0 + d2) + 0.077 * torch.eye(n, device=dev)
B = torch.randn(n, t, device=dev, generator=g)
ref = torch.linalg.solve(K.double(), B.double()).float()
for lag in (0, 2, 1):
    with settings.cg_convergence_lag(lag):
        x, info = linear_cg(lambda v: K @ v, B, tolerance=1e-3, max_iter=1000, return_info=True)
    print("lag", lag, info, "rel err", float((x - ref).norm() / ref.norm()), "finite", bool(torch.isfinite(x).all()))
