#!/bin/bash
# round 2, GPU call J: K > 1 kernel after the elect-based arrives: parity tests, timings, cost breakdown (diagnostic builds), timeline
mkdir -p gpurun_out
O=gpurun_out/tcd_j.txt; : > $O
timeout 600 python -m pytest tests/test_sym_tc_gpu.py -m gpu -q -x 2>&1 | tail -3 >> $O
for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
  echo "=== base shape=$shape" >> $O
  timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
done
for v in diag1 diag3 diag7 diag8 diag15 diag47; do
  for shape in "100000 1 20" "100000 20 5"; do
    echo "=== variant=$v shape=$shape" >> $O
    RPGP_LIB=build/librpgp_$v.so timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
  done
done
for shape in "100000 1 20" "100000 20 5"; do
  echo "=== variant=stamps shape=$shape" >> $O
  RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_stamps.so timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -34 >> $O
done
cat $O
