#!/bin/bash
# round 2: the cosine base kernel (base 3) on the GPU: base-kernel tests, symmetric-kernel tests, smoke, timings (hard limits)
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_base_kernels_gpu.py tests/test_sym_tc_gpu.py tests/test_cabi_gpu.py -m gpu -q > gpurun_out/pytest_cos.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_cos.txt
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/pytest_cos.txt | head -40
grep -E "^E  " gpurun_out/pytest_cos.txt | head -30
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout -s KILL 200 python tools/next_rows_bench.py base > gpurun_out/next_rows_r02_base.txt 2>&1
cut -c1-200 gpurun_out/next_rows_r02_base.txt
