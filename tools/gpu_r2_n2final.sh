#!/bin/bash
# round 2: configs[4] on two GPUs with the shipped K > 1 kernel (hard limit)
mkdir -p gpurun_out
for wl in cfg5a; do
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --mll-workload none > gpurun_out/bench_r02_final_${wl}_n2.json 2> gpurun_out/bench_r02_final_${wl}_n2.err
  echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02_final_${wl}_n2.json").read().strip().splitlines()[-1])
    print("${wl} N=2", "ms/step %.1f" % d["ms_per_step"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d["parity"] and (d["parity"]["norm_rel"], d["parity"]["max_row_rel"], d["parity"]["ok"]))
except Exception as e:
    print("${wl} failed", e)
PY
done
