#!/bin/bash
# A/B of builds of sym_tc5.cu (build/librpgp_<name>.so, made with -DT5_QUNROLL / -DT5_REGS_* / -DT5_DIAG); timing only
LIB=randomly-projected-additive-gps_b200/rpgp/librpgp.so
cp $LIB /tmp/keep.so
echo "default build, polynomial pairs 0..3:"; for np in 0 1 2 3; do RPGP_SYM_POLY_PAIRS=$np N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1; done
for v in "$@"; do
  cp build/librpgp_$v.so $LIB
  echo "variant $v:"; N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1
done
cp /tmp/keep.so $LIB
