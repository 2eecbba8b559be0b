"""Small driver for ncu: a few launches of the symmetric tensor-core forward at n=50k, J=20, t=11."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "randomly-projected-additive-gps_b200"))
import torch
from rpgp import _lib
dev = torch.device("cuda:0")
n, J, t = int(os.environ.get("N", "50000")), 20, 11
lay = _lib.plan_layout(J, 1)
g = torch.Generator(device=dev); g.manual_seed(0)
Z = torch.randn(n, J, device=dev, generator=g) * 4.5
zp = _lib.pack_coords(Z, lay)
nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
V = torch.randn(n, t, device=dev, generator=g)
for _ in range(3):
    _lib.mvm_sym(zp, lay, nlc, V)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); _lib.mvm_sym(zp, lay, nlc, V); e1.record(); torch.cuda.synchronize()
print("n=%d: %.3f ms, %.3f T pair-evals/s" % (n, e0.elapsed_time(e1), n * n * J / e0.elapsed_time(e1) / 1e9))
