#!/bin/bash
# round 2, GPU call M: K > 1 kernel with setmaxnreg (96 registers for the arithmetic warps) + opaque indices; timing probes of the tile tail
mkdir -p gpurun_out
O=gpurun_out/tcd_m.txt; : > $O
timeout 900 python -m pytest tests/test_sym_tc_gpu.py tests/test_base_kernels_gpu.py -m gpu -q 2>&1 | tail -40 >> $O
for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
  echo "=== base shape=$shape" >> $O
  timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
done
for v in probe1 probe2 probe3 probe4 probe7; do
  for shape in "100000 20 5" "100000 1 20"; do
    echo "=== variant=$v shape=$shape" >> $O
    RPGP_LIB=build/librpgp_$v.so timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
  done
done
echo "=== variant=stamps shape=100000 20 5" >> $O
RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_stamps.so timeout 120 python tools/tcd_check.py time 100000 20 5 2>&1 | tail -43 | head -16 >> $O
cat $O
