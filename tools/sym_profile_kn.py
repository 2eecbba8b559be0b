"""Timing driver for the symmetric kernel on a K > 1 layout: N rows, J groups of K coordinates, t = 11."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "randomly-projected-additive-gps_b200"))
import torch
from rpgp import _lib
dev = torch.device("cuda:0")
n, J, K, t = int(os.environ.get("N", "100000")), int(os.environ.get("J", "1")), int(os.environ.get("K", "20")), 11
lay = _lib.plan_layout(J, K)
g = torch.Generator(device=dev); g.manual_seed(0)
Z = torch.randn(n, J * K, device=dev, generator=g) * 1.5
zp = _lib.pack_coords(Z, lay)
nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
V = torch.randn(n, t, device=dev, generator=g)
for _ in range(3):
    _lib.mvm_sym(zp, lay, nlc, V)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); _lib.mvm_sym(zp, lay, nlc, V); e1.record(); torch.cuda.synchronize()
print("n=%d J=%d K=%d (CP=%d KP=%d G=%d chunks=%d): %.3f ms" % (n, J, K, lay.CP, lay.KP, lay.G, lay.nchunks, e0.elapsed_time(e1)))
