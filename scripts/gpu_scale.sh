#!/bin/bash
# usage: gpu_scale.sh N  -- the bench under torchrun at N ranks (headline workload + BASELINE configs 3 and 4)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read())
print({k:d[k] for k in ("value","ms_per_step","n_gpus","dp_check","rec_loss_mean_over_ranks")}, "e2e", d["e2e"]["value"])
a=d["allreduce"]; print("allreduce", {k:a[k] for k in a if k!="note"})
for k,v in d.get("configs",{}).items():
    a=v.get("allreduce") or {}
    print(k, {x:v.get(x) for x in ("value","ms_per_step","dp_check")}, "e2e", (v.get("e2e") or {}).get("value"), {x:a[x] for x in a if x!="note"})
PY
