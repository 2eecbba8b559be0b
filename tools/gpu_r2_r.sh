#!/bin/bash
# round 2, GPU call R: probe -- S stores compiled in but never executed (exponentials stay live)
mkdir -p gpurun_out
O=gpurun_out/tcd_r.txt; : > $O
for v in base nosts; do
  L=randomly-projected-additive-gps_b200/rpgp/librpgp.so; [ $v != base ] && L=build/librpgp_$v.so
  for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
    echo "=== $v shape=$shape" >> $O
    RPGP_LIB=$L timeout -s KILL 50 python tools/tcd_check.py time $shape > gpurun_out/q.tmp 2>&1; echo "rc=$?" >> $O; tail -1 gpurun_out/q.tmp >> $O
  done
done
echo "=== variant=nostsstamps shape=100000 20 5" >> $O
RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_nostsstamps.so timeout -s KILL 50 python tools/tcd_check.py time 100000 20 5 2>&1 | tail -43 | head -10 >> $O
cat $O
