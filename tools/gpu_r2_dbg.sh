#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/dbg_imq.txt
for cfg in "-1 0" "0 0" "-1 1" "1 0"; do set -- $cfg
  echo "=== LAG=$1 CUDA_LAUNCH_BLOCKING=$2" >> gpurun_out/dbg_imq.txt
  CUDA_LAUNCH_BLOCKING=$2 LAG=$1 timeout -s KILL 150 python tools/debug_imq.py 2>&1 | tail -5 >> gpurun_out/dbg_imq.txt
done
cat gpurun_out/dbg_imq.txt
