#!/bin/bash
# round 2, GPU call F: K > 1 kernel with the two teams kept half a tile out of phase
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sym_tc_gpu.py -x -q > gpurun_out/pytest_f.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_f.txt
{
for v in "" pp0; do
  if [ -n "$v" ]; then export RPGP_LIB=$PWD/build/librpgp_$v.so; else unset RPGP_LIB; fi
  for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
    echo -n "variant=${v:-default} "; timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1
  done
done
unset RPGP_LIB
} > gpurun_out/times_f.txt 2>&1
for shape in "100000 20 5"; do
  echo "=== variant=stamps shape=$shape"
  RPGP_LIB=$PWD/build/librpgp_stamps.so RPGP_TCD_DBG=1 timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -31
done > gpurun_out/stamps_f.txt 2>&1
tail -4 gpurun_out/pytest_f.txt; cat gpurun_out/times_f.txt; cat gpurun_out/stamps_f.txt
