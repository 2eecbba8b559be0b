#!/bin/bash
# first full GPU pass: smoke, bench (cfg4 + cfg2), reference arm, ncu launch list + full capture of the forward kernel
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 3000 gpurun_out/bench_cfg4.json; tail -5 gpurun_out/bench_cfg4.err
python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 1500 gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_cfg4.json 2>&1; tail -c 800 gpurun_out/bench_ref_cfg4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --workload cfg2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mvm_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd_cfg2 \
    python bench.py --workload cfg2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
