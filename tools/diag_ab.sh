#!/bin/bash
# A/B of diagnostic builds of the symmetric kernel (bit 1: no epilogue atomics, 2: no MMAs, 4: no S stores, 8: no TMEM loads)
cp randomly-projected-additive-gps_b200/rpgp/librpgp.so /tmp/keep.so
echo "baseline:"; N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1
for d in 1 2 4 7 15; do
  cp build/librpgp_d$d.so randomly-projected-additive-gps_b200/rpgp/librpgp.so
  echo "T3_DIAG=$d:"; N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1
done
cp /tmp/keep.so randomly-projected-additive-gps_b200/rpgp/librpgp.so
