#!/bin/bash
# round 2, final GPU call: whole GPU suite, smoke, default bench (cfg4), single-GPU lines of configs[4] with the shipped K > 1 kernel.
# Every step under a hard limit (a hung kernel must not eat the budget).
mkdir -p gpurun_out
timeout -s KILL 1100 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_final.txt
timeout -s KILL 600 python bench.py > gpurun_out/bench_r02_final_cfg4_n1.json 2> gpurun_out/bench_final_cfg4.err
echo "bench rc=$?" >> gpurun_out/bench_final_cfg4.err
for wl in cfg5a cfg5b; do
  timeout -s KILL 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --mll-workload none > gpurun_out/bench_r02_final_${wl}_n1.json 2> gpurun_out/bench_final_${wl}.err
  echo "$wl rc=$?" >> gpurun_out/bench_final_${wl}.err
done
tail -6 gpurun_out/pytest_final.txt; tail -2 gpurun_out/smoke_final.txt; tail -1 gpurun_out/bench_final_cfg4.err; tail -1 gpurun_out/bench_final_cfg5a.err; tail -1 gpurun_out/bench_final_cfg5b.err
for f in gpurun_out/bench_r02_final_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['ms_per_step'], d['roofline']['frac'], d['parity']['max_row_rel'], d['parity']['ok'], d.get('e2e',{}).get('ms_per_step'))
except Exception as e: print(sys.argv[1], 'unreadable', e)
PY
done
