#!/bin/bash
# round 2, GPU call Q: deferred tile tail vs not, registers 96 / 56 / 40 from a launch at 80 -- every run under a hard limit
mkdir -p gpurun_out
O=gpurun_out/tcd_q.txt; : > $O
for v in d0 base; do
  L=randomly-projected-additive-gps_b200/rpgp/librpgp.so; [ $v != base ] && L=build/librpgp_$v.so
  for shape in "100000 8 6" "100000 20 5" "100000 1 20"; do
    echo "=== $v shape=$shape" >> $O
    RPGP_LIB=$L timeout -s KILL 50 python tools/tcd_check.py time $shape > gpurun_out/q.tmp 2>&1; echo "rc=$?" >> $O; tail -1 gpurun_out/q.tmp >> $O
  done
done
timeout -s KILL 240 python -m pytest tests/test_sym_tc_gpu.py -m gpu -q -x > gpurun_out/q.tmp 2>&1; echo "pytest rc=$?" >> $O; tail -3 gpurun_out/q.tmp >> $O
cat $O
