"""SIMT forward vs symmetric tensor-core forward across n (J=20, t=11): where should the operator switch?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "randomly-projected-additive-gps_b200"))
import torch
from rpgp import _lib
dev = torch.device("cuda:0")
J, t = 20, 11
lay = _lib.plan_layout(J, 1)
def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1024, 2000, 4000, 8000, 16000, 32000, 64000):
    g = torch.Generator(device=dev); g.manual_seed(0)
    Z = torch.randn(n, J, device=dev, generator=g) * 2.0
    zp = _lib.pack_coords(Z, lay); nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
    V = torch.randn(n, t, device=dev, generator=g)
    reps = max(3, min(200, int(2e9 / (n * n))))
    a = timeit(lambda: _lib.mvm_fwd(zp, zp, lay, nlc, V), reps)
    b = timeit(lambda: _lib.mvm_sym(zp, lay, nlc, V), reps)
    print(f"n={n:6d}: SIMT {a:8.3f} ms   sym-TC {b:8.3f} ms   ratio {a/b:.2f}", flush=True)
