#!/bin/bash
# round 2, GPU call G: projection on tcgen05 + its VJP: parity, model tests, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_project_gpu.py -x -q > gpurun_out/pytest_g1.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_g1.txt
timeout 1200 python -m pytest tests/test_cabi_gpu.py tests/test_model_gpu.py tests/test_base_kernels_gpu.py -x -q > gpurun_out/pytest_g2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_g2.txt
timeout 300 python tools/project_time.py > gpurun_out/project_time_g.txt 2>&1
tail -15 gpurun_out/pytest_g1.txt; tail -15 gpurun_out/pytest_g2.txt; cat gpurun_out/project_time_g.txt
