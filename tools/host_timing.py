"""Phases of the host-buffer entry point rpgp_kmv_host_f32 (RPGP_HOST_TIMING=1) at a bench workload: python tools/host_timing.py cfg5b"""
import os
import sys
import time

os.environ["RPGP_HOST_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "randomly-projected-additive-gps_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
from rpgp import _lib  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5b"]
X, W, inv_ell, c, V = bench.make_inputs(w)
for rep in range(2):
    t0 = time.perf_counter()
    _lib.kmv_host(X.numpy(), None, W.numpy(), w["J"], w["K"], inv_ell.numpy(), None, c.numpy(), V.numpy(), diag_add=1.0)
    print("call %d: %.1f ms wall" % (rep, 1e3 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
