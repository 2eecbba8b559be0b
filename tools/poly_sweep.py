"""Sweep RPGP_POLY_PAIRS (packed projection pairs per (i,i') evaluated by the FMA-pipe polynomial exp2): accuracy vs the
FP64 oracle and time at cfg2 scale.  One process per setting (the library reads the variable once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(%r, "randomly-projected-additive-gps_b200")); sys.path.insert(0, %r)
from rpgp import _lib
from oracle import rpgp_oracle as orc
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
res = []
for spread in (1.0, 4.5, 9.5):
    Z1 = (rng.randn(512, 20) * spread).astype(np.float32); Z2 = (rng.randn(6000, 20) * spread).astype(np.float32)
    c = (rng.rand(20) * 2 + 0.01).astype(np.float32); V = rng.randn(6000, 11).astype(np.float32)
    lay = _lib.plan_layout(20, 1)
    z1 = _lib.pack_coords(torch.from_numpy(Z1).to(dev), lay); z2 = _lib.pack_coords(torch.from_numpy(Z2).to(dev), lay)
    nlc = _lib.pack_log2c(torch.from_numpy(c).to(dev), lay)
    out = _lib.mvm_fwd(z1, z2, lay, nlc, torch.from_numpy(V).to(dev)).cpu().numpy()
    ref = orc.kmv(Z1, Z2, c, 20, 1, V)
    res.append(np.linalg.norm(out - ref) / np.linalg.norm(ref))
n = 100_000
g = torch.Generator(device=dev); g.manual_seed(0)
Z = torch.randn(n, 20, device=dev, generator=g) * 4.5
zp = _lib.pack_coords(Z, lay); nlc = _lib.pack_log2c(torch.full((20,), 0.03, device=dev), lay)
V = torch.randn(n, 11, device=dev, generator=g)
for _ in range(3): _lib.mvm_fwd(zp, zp, lay, nlc, V)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): _lib.mvm_fwd(zp, zp, lay, nlc, V)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("RPGP_POLY_PAIRS=%%s  rel.err %%s  cfg2 %%.3f ms  %%.3f T pair-evals/s (%%.1f%%%% of 4.64e12 MUFU roof)" %% (
    os.environ.get("RPGP_POLY_PAIRS", "0"), ["%%.2e" %% r for r in res], ms, n * n * 20 / ms / 1e9, n * n * 20 / ms / 1e9 / 4.64 * 100))
''' % (ROOT, ROOT)
for npairs in sys.argv[1:] or ["0", "1", "2", "3", "4", "5"]:
    env = dict(os.environ, RPGP_POLY_PAIRS=npairs)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
