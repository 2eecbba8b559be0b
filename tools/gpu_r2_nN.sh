#!/bin/bash
# round 2, N GPUs (argument 1): bench lines at N ranks for the workloads given as the remaining arguments
N=$1; shift
mkdir -p gpurun_out
for wl in "$@"; do
  extra="--no-cpu-baseline --mll-workload none"
  if [ "$wl" = "cfg4" ]; then extra=""; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --steps 3 --warmup 3 $extra > gpurun_out/bench_r02_${wl}_n$N.json 2> gpurun_out/bench_r02_${wl}_n$N.err
  echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02_${wl}_n$N.json").read().strip().splitlines()[-1])
    print("${wl} N=$N", "ms/step %.1f" % d["ms_per_step"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d["parity"] and (d["parity"]["norm_rel"], d["parity"]["max_row_rel"], d["parity"]["ok"]), "mll", d.get("mll_step") and d["mll_step"]["ms_per_step"])
except Exception as e:
    print("${wl} failed", e)
PY
done
