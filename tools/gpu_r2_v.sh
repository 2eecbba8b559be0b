#!/bin/bash
# round 2, GPU call V: timeline of the K > 1 kernel after the lean hand-off
mkdir -p gpurun_out
O=gpurun_out/tcd_v.txt; : > $O
for shape in "100000 20 5" "100000 1 20"; do
echo "=== variant=stamps shape=$shape" >> $O
RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_stamps.so timeout -s KILL 50 python tools/tcd_check.py time $shape 2>&1 | tail -43 | head -12 >> $O
done
cat $O
