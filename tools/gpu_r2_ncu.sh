#!/bin/bash
# round 2: ncu --set full captures of the three kernels that changed (one GPU)
mkdir -p gpurun_out
N=100000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mvm_sym_tc5 -s 2 -c 1 -f -o gpurun_out/prof_r02_sym5_cfg2 python tools/sym_profile.py > gpurun_out/ncu_sym5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mvm_sym_tcd -s 2 -c 1 -f -o gpurun_out/prof_r02_tcd_cfg5b python tools/tcd_check.py time 100000 20 5 > gpurun_out/ncu_tcd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_tc -s 2 -c 1 -f -o gpurun_out/prof_r02_project_cfg4 python tools/project_time.py > gpurun_out/ncu_project.log 2>&1
ls -la gpurun_out/*.ncu-rep
# single-GPU lines of the other configurations with the round-2 kernels
for wl in cfg3 cfg5a cfg5b cfg2; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --mll-workload none > gpurun_out/bench_r02_${wl}_n1.json 2> gpurun_out/bench_r02_${wl}_n1.err
  echo "$wl rc=$?"
done
