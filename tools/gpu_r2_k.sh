#!/bin/bash
# round 2, GPU call K: K > 1 kernel: fine-grained timeline (shared-memory stamps), one-time stagger of team 1
mkdir -p gpurun_out
O=gpurun_out/tcd_k.txt; : > $O
for v in stag1000 stag1700 stag2500; do
  for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
    echo "=== variant=$v shape=$shape" >> $O
    RPGP_LIB=build/librpgp_$v.so timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
  done
done
for v in stamps stampstag; do
for shape in "100000 20 5" "100000 1 20"; do
  echo "=== variant=$v shape=$shape" >> $O
  RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_$v.so timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -43 >> $O
done
done
cat $O
