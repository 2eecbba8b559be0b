#!/bin/bash
# round 2, final GPU call (second pass): whole GPU suite, smoke, default bench line -- every step under a hard limit
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_final.txt
timeout -s KILL 600 python bench.py > gpurun_out/bench_r02_final_cfg4_n1.json 2> gpurun_out/bench_final_cfg4.err
echo "bench rc=$?" >> gpurun_out/bench_final_cfg4.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final_cfg4.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --mll-workload none > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/pytest_final.txt; tail -2 gpurun_out/smoke_final.txt; tail -1 gpurun_out/bench_final_cfg4.err; wc -l gpurun_out/launches_r02_final_cfg4.csv
