#!/bin/bash
# diagnostics sweep of csrc/sym_tcd.cu (RPGP_TCD_DIAG bits: 1 no column atomics, 2 no column MMAs, 4 no row MMAs, 8 no distance MMAs,
# 16 no exponentials, 32 no D2 reads)
for shape in "100000 1 20" "100000 8 6"; do
  for d in 0 1 2 4 8 6 14 16 30 33 63; do
    echo -n "diag=$d "; RPGP_TCD_DIAG=$d timeout 60 python tools/tcd_check.py time $shape 2>&1 | tail -1
  done
done
