#!/bin/bash
# diagnostics sweep of csrc/sym_tcd.cu: rebuilds the library with -DTCD_DIAG=<bits> (1 no column atomics, 2 no column MMAs, 4 no row
# MMAs, 8 no distance MMAs, 32 no D2 reads) and times two shapes each; restores the normal build at the end.
# (profiles/tcd_diag_r01.txt was taken with the same switches as run-time flags, before they became compile-time.)
set -e
cd "$(dirname "$0")/.."
CS=randomly-projected-additive-gps_b200/csrc
for d in 0 1 2 4 8 6 14 33 47; do
  touch $CS/sym_tcd.cu
  make -s -C $CS NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -DTCD_DIAG=$d" > /dev/null
  for shape in "100000 1 20" "100000 8 6"; do
    echo -n "diag=$d "; timeout 60 python tools/tcd_check.py time $shape 2>&1 | tail -1
  done
done
touch $CS/sym_tcd.cu; make -s -C $CS > /dev/null
