#!/bin/bash
# projection kernel diagnostics: which stage limits the copy
for v in "" pd1 pd2 pd4 pd8 pd3 pd15; do
  if [ -n "$v" ]; then export RPGP_LIB=$PWD/build/librpgp_$v.so; else unset RPGP_LIB; fi
  echo -n "variant=${v:-default} "; timeout 120 python tools/project_time.py 2>&1 | grep "cfg4   project (packed)"
done
