#!/bin/bash
for sp in 1 2 4 8 16; do echo "RPGP_SYM_SPLITS=$sp"; RPGP_SYM_SPLITS=$sp N=1000000 timeout 100 python tools/sym_profile.py 2>&1 | tail -1; done
