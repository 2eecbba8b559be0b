#!/bin/bash
# round 2, GPU call I: whole GPU suite, smoke, default bench (cfg4), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_i.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_i.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_i.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_i_cfg4.json 2> gpurun_out/bench_i_cfg4.err
echo "bench rc=$?" >> gpurun_out/bench_i_cfg4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_cfg4.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --mll-workload none > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/pytest_i.txt; cat gpurun_out/smoke_i.txt | tail -2; tail -2 gpurun_out/bench_i_cfg4.err; wc -l gpurun_out/launches_r02_cfg4.csv
