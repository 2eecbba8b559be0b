#!/bin/bash
# A/B: default library vs q-loop unrolled x2
N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1
cp randomly-projected-additive-gps_b200/rpgp/librpgp.so /tmp/librpgp_keep.so
cp build/alt/librpgp_u2.so randomly-projected-additive-gps_b200/rpgp/librpgp.so
echo "unroll 2:"; N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1
timeout 100 python -m pytest tests/test_sym_tc_gpu.py -x -q -m gpu 2>&1 | tail -1
cp /tmp/librpgp_keep.so randomly-projected-additive-gps_b200/rpgp/librpgp.so
