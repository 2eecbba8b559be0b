#!/bin/bash
# round 2, GPU call A: MMA issue-rate microbenchmark (lane==0 vs elect.sync), kernel timings after the elect change, GPU tests, default bench
mkdir -p gpurun_out
./build/umma_rate2 > gpurun_out/umma_rate2.txt 2>&1
{
python tools/tcd_check.py time 100000 20 5
python tools/tcd_check.py time 100000 1 20
python tools/tcd_check.py time 100000 8 6
N=100000 python tools/sym_profile.py
} > gpurun_out/times_a.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_a.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a_cfg4.json 2> gpurun_out/bench_a_cfg4.err
echo "bench rc=$?" >> gpurun_out/bench_a_cfg4.err
tail -5 gpurun_out/pytest_a.txt; cat gpurun_out/times_a.txt; tail -3 gpurun_out/bench_a_cfg4.err
