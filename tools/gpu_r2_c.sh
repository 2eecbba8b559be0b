#!/bin/bash
# round 2, GPU call C: in-kernel timeline of the K > 1 kernel (clock64 stamps of one CTA), full and skeleton (TCD_DIAG=47) builds
mkdir -p gpurun_out
for v in stamps stamps47; do
  for shape in "100000 20 5" "100000 1 20"; do
    echo "=== variant=$v shape=$shape"
    RPGP_LIB=$PWD/build/librpgp_$v.so RPGP_TCD_DBG=1 timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -32
  done
done > gpurun_out/stamps_c.txt 2>&1
cat gpurun_out/stamps_c.txt
