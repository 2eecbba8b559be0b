#!/bin/bash
# round 2, GPU call N: conflict-free S stores in both symmetric kernels; the failing inverse-MQ training test against the committed build
mkdir -p gpurun_out
O=gpurun_out/tcd_n.txt; : > $O
timeout 900 python -m pytest tests/test_sym_tc_gpu.py tests/test_base_kernels_gpu.py -m gpu -q 2>&1 | tail -4 >> $O
echo "=== inverse-MQ training test with the committed build" >> $O
RPGP_LIB=build/librpgp_head.so timeout 600 python -m pytest tests/test_base_kernels_gpu.py -m gpu -q -k inverse_multiquadric 2>&1 | tail -4 >> $O
for shape in "100000 20 5" "100000 1 20" "100000 8 6"; do
  echo "=== base shape=$shape" >> $O
  timeout 120 python tools/tcd_check.py time $shape 2>&1 | tail -1 >> $O
done
for w in cfg2 cfg4; do
  echo "=== bench $w" >> $O
  timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline --mll-workload none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_row_rel'], d['parity']['ok'])" >> $O 2>&1
done
echo "=== variant=stamps shape=100000 20 5" >> $O
RPGP_TCD_DBG=1 RPGP_LIB=build/librpgp_stamps.so timeout 120 python tools/tcd_check.py time 100000 20 5 2>&1 | tail -43 | head -12 >> $O
cat $O
