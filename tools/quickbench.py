import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "randomly-projected-additive-gps_b200"))
import torch
from rpgp import _lib
dev = torch.device("cuda:0")
def run(n, m, J, t, spread, reps=3):
    lay = _lib.plan_layout(J, 1)
    g = torch.Generator(device=dev); g.manual_seed(0)
    Z = torch.randn(n, J, device=dev, generator=g) * spread
    zp = _lib.pack_coords(Z, lay)
    nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
    V = torch.randn(n, t, device=dev, generator=g)
    rr = (0, m)
    for _ in range(2): _lib.mvm_fwd(zp, zp, lay, nlc, V, row_range=rr)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.mvm_fwd(zp, zp, lay, nlc, V, row_range=rr)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pe = m * n * J / (ms * 1e-3)
    print(f"n={n} m={m} J={J} t={t} spread={spread}: {ms:.3f} ms  {pe/1e12:.3f} T pair-evals/s  ({pe/4.65e12*100:.1f}% of 4.65e12 MUFU roof)", flush=True)
print(_lib.measure_peaks())
run(2000, 2000, 20, 11, 1.0, 20)
run(100_000, 100_000, 20, 11, 1.0)
run(100_000, 100_000, 20, 11, 4.5)
run(100_000, 100_000, 20, 1, 1.0)
run(100_000, 100_000, 20, 16, 1.0)
run(400_000, 100_000, 26, 16, 1.4)
run(1_000_000, 65536, 20, 11, 9.5)

def run_bwd(n, J, t, spread, reps=3):
    lay = _lib.plan_layout(J, 1)
    g = torch.Generator(device=dev); g.manual_seed(0)
    Z = torch.randn(n, J, device=dev, generator=g) * spread
    zp = _lib.pack_coords(Z, lay)
    nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
    L = torch.randn(n, t, device=dev, generator=g); R = torch.randn(n, t, device=dev, generator=g)
    for _ in range(2): _lib.quad_bwd(zp, zp, lay, nlc, L, R, symmetric=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.quad_bwd(zp, zp, lay, nlc, L, R, symmetric=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pe = n * n * J / (ms * 1e-3)
    print(f"BWD sym n={n} J={J} t={t}: {ms:.3f} ms  {pe/1e12:.3f} T pair-evals/s  ({pe/4.65e12*100:.1f}% of MUFU roof)", flush=True)
run_bwd(100_000, 20, 11, 1.0)
run_bwd(100_000, 20, 11, 4.5)
run_bwd(100_000, 26, 16, 1.4)

def run_sym(n, J, t, spread, reps=3):
    lay = _lib.plan_layout(J, 1)
    g = torch.Generator(device=dev); g.manual_seed(0)
    Z = torch.randn(n, J, device=dev, generator=g) * spread
    zp = _lib.pack_coords(Z, lay)
    nlc = _lib.pack_log2c(torch.full((J,), 0.03, device=dev), lay)
    V = torch.randn(n, t, device=dev, generator=g)
    for _ in range(2): _lib.mvm_sym(zp, lay, nlc, V)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.mvm_sym(zp, lay, nlc, V)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pe = n * n * J / (ms * 1e-3)
    print(f"SYM-TC n={n} J={J} t={t} spread={spread}: {ms:.3f} ms  {pe/1e12:.3f} T pair-evals/s  ({pe/4.65e12*100:.1f}% of the MUFU roof of the non-symmetric algorithm)", flush=True)
import os
if os.environ.get("RPGP_SYM_BENCH", "1") == "1":
    run_sym(100_000, 20, 11, 1.0)
    run_sym(100_000, 20, 11, 4.5)
    run_sym(200_000, 20, 11, 4.5, reps=2)
