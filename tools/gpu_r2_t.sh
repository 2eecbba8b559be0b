#!/bin/bash
# round 2, GPU call T/U: lean batch hand-off, loop parameters through volatile shared loads (hard limits on every run)
mkdir -p gpurun_out
O=gpurun_out/tcd_y.txt; : > $O
for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
  echo "=== base shape=$shape" >> $O
  timeout -s KILL 50 python tools/tcd_check.py time $shape > gpurun_out/q.tmp 2>&1; echo "rc=$?" >> $O; tail -1 gpurun_out/q.tmp >> $O
done
timeout -s KILL 240 python -m pytest tests/test_sym_tc_gpu.py -m gpu -q -x > gpurun_out/q.tmp 2>&1; echo "pytest rc=$?" >> $O; tail -3 gpurun_out/q.tmp >> $O
cat $O
