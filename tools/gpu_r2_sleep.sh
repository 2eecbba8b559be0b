#!/bin/bash
# round 2: back-off of the helper warps' barrier polls in the K > 1 kernel (run-time switch RPGP_TCD_SLEEP, ns)
mkdir -p gpurun_out
O=gpurun_out/tcd_sleep.txt; : > $O
for ns in 64 0 20 200; do
  for shape in "100000 20 5" "100000 1 20"; do
    echo "=== sleep=$ns shape=$shape" >> $O
    RPGP_TCD_SLEEP=$ns timeout -s KILL 50 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
  done
done
cat $O
