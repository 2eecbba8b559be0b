#!/bin/bash
# round 2, GPU call B: A-in-TMEM probe; K > 1 kernel with one distance issuer per team (3 / 2 D0 buffers) and its diagnostic builds;
# parity of the symmetric kernels; projection kernel timing
mkdir -p gpurun_out
./build/umma_probe4 > gpurun_out/umma_probe4.txt 2>&1
{
for v in "" nbuf2 d8 d6 d14 d47; do
  if [ -n "$v" ]; then export RPGP_LIB=$PWD/build/librpgp_$v.so; else unset RPGP_LIB; fi
  for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
    echo -n "variant=${v:-default} "; timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1
  done
done
unset RPGP_LIB
} > gpurun_out/times_b.txt 2>&1
timeout 900 python -m pytest tests/test_sym_tc_gpu.py -x -q > gpurun_out/pytest_b.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_b.txt
python - > gpurun_out/project_time.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, "randomly-projected-additive-gps_b200")
from rpgp import _lib
dev = torch.device("cuda:0")
for (n, d, J, K) in [(1_000_000, 90, 20, 1), (1_000_000, 90, 20, 5), (100_000, 20, 20, 1), (400_000, 26, 26, 1)]:
    lay = _lib.plan_layout(J, K)
    X = torch.randn(n, d, device=dev); W = torch.randn(J * K, d, device=dev); ell = torch.ones(d, device=dev)
    for _ in range(3): zp = _lib.project(X, W, ell, None, lay)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): zp = _lib.project(X, W, ell, None, lay)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = 4.0 * (n * d + zp.numel())
    print("project n=%d d=%d J=%d K=%d: %.3f ms, %.0f GB/s algorithmic (read X + write packed Z^)" % (n, d, J, K, ms, byts / ms / 1e6))
PY
cat gpurun_out/umma_probe4.txt gpurun_out/times_b.txt gpurun_out/project_time.txt; tail -4 gpurun_out/pytest_b.txt
