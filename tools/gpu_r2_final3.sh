#!/bin/bash
# round 2, final GPU call (third pass, after the cosine base kernel): whole GPU suite, smoke, base-kernel timings, default bench line
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_final.txt
timeout -s KILL 200 python tools/next_rows_bench.py base > gpurun_out/next_rows_r02_base.txt 2>&1
timeout -s KILL 600 python bench.py > gpurun_out/bench_r02_final_cfg4_n1.json 2> gpurun_out/bench_final_cfg4.err
echo "bench rc=$?" >> gpurun_out/bench_final_cfg4.err
tail -4 gpurun_out/pytest_final.txt; tail -2 gpurun_out/smoke_final.txt; tail -1 gpurun_out/bench_final_cfg4.err; cat gpurun_out/next_rows_r02_base.txt | cut -c1-200
