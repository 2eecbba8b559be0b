#!/bin/bash
# sweep the number of polynomial-exp2 pairs in the symmetric kernel (cfg2 scale) + correctness
timeout 200 python -m pytest tests/test_sym_tc_gpu.py -x -q -m gpu 2>&1 | tail -2
for np in 0 1 2; do echo "RPGP_SYM_POLY_PAIRS=$np"; RPGP_SYM_POLY_PAIRS=$np N=100000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1; done
