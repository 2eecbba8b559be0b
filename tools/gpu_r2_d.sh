#!/bin/bash
# round 2, GPU call D: restructured K > 1 kernel (TDONE wait before the S stores, proxy fence on the issuer, dedicated copy warp with
# early B-image release): parity, timings, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sym_tc_gpu.py -x -q > gpurun_out/pytest_d.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_d.txt
{
for v in "" fence0 d47; do
  if [ -n "$v" ]; then export RPGP_LIB=$PWD/build/librpgp_$v.so; else unset RPGP_LIB; fi
  for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
    echo -n "variant=${v:-default} "; timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1
  done
done
unset RPGP_LIB
} > gpurun_out/times_d.txt 2>&1
for v in stamps stamps47; do
  for shape in "100000 20 5" "100000 1 20"; do
    echo "=== variant=$v shape=$shape"
    RPGP_LIB=$PWD/build/librpgp_$v.so RPGP_TCD_DBG=1 timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -31
  done
done > gpurun_out/stamps_d.txt 2>&1
tail -4 gpurun_out/pytest_d.txt; cat gpurun_out/times_d.txt; cat gpurun_out/stamps_d.txt
