#!/bin/bash
# round 2, 2 GPUs: NCCL parity test of the row-partitioned model, bench at N = 2 (cfg4 with parity after the all-reduce, cfg3, cfg5b)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -q > gpurun_out/pytest_n2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_n2.txt
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 $2 > gpurun_out/bench_r02_$1_n2.json 2> gpurun_out/bench_r02_$1_n2.err
  echo "$1 rc=$?"; tail -c 400 gpurun_out/bench_r02_$1_n2.json | head -c 10 > /dev/null
}
run cfg4 "--steps 3 --warmup 3"
run cfg3 "--workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --mll-workload none"
run cfg5b "--workload cfg5b --steps 3 --warmup 3 --no-cpu-baseline --mll-workload none"
tail -3 gpurun_out/pytest_n2.txt
for f in cfg4 cfg3 cfg5b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02_${f}_n2.json").read().strip().splitlines()[-1])
    print("${f}", "ms/step %.1f" % d["ms_per_step"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d["parity"] and (d["parity"]["norm_rel"], d["parity"]["max_row_rel"], d["parity"]["ok"]), "mll", d.get("mll_step") and d["mll_step"]["ms_per_step"])
except Exception as e:
    print("${f} failed", e)
PY
done
