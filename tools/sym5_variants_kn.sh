#!/bin/bash
LIB=randomly-projected-additive-gps_b200/rpgp/librpgp.so
cp $LIB /tmp/keep.so
for cfg in "1 20" "20 5"; do set -- $cfg
  echo "default:"; J=$1 K=$2 timeout 100 python tools/sym_profile_kn.py 2>&1 | tail -1
  for v in d1 d6 d7; do cp build/librpgp_$v.so $LIB; echo "variant $v:"; J=$1 K=$2 timeout 100 python tools/sym_profile_kn.py 2>&1 | tail -1; cp /tmp/keep.so $LIB; done
done
