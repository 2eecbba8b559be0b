#!/bin/bash
export RPGP_SYM_IMPL=4
timeout 200 python -m pytest tests/test_sym_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
for np in 0 1; do echo "v4 RPGP_SYM_POLY_PAIRS=$np"; RPGP_SYM_POLY_PAIRS=$np N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1; done
