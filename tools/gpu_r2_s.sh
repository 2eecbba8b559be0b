#!/bin/bash
# round 2, GPU call S: ncu --set full (with source) of the K > 1 kernel after the S2R / setmaxnreg changes
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:mvm_sym_tcd -s 2 -c 1 -f -o gpurun_out/prof_r02c_tcd_cfg5b python tools/tcd_check.py time 100000 20 5 > gpurun_out/ncu_tcd_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_tcd_c.log
