#!/bin/bash
# 2-rank run on hardware: correctness of the in-library bucketed all-reduce (scripts/dp_two_ranks.py), then the bench under
# torchrun (headline workload + BASELINE configs 3 and 4, all-reduce accounting, replica check).
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29520 scripts/dp_two_ranks.py > gpurun_out/dp_two_ranks.json 2> gpurun_out/dp_two_ranks.err; echo "dp_two_ranks rc=$?"; cut -c1-1500 gpurun_out/dp_two_ranks.json; tail -5 gpurun_out/dp_two_ranks.err
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read())
print({k:d[k] for k in ("value","ms_per_step","n_gpus","dp_check","rec_loss_mean_over_ranks")}, "e2e", d["e2e"]["value"])
print("allreduce", d["allreduce"])
for k,v in d.get("configs",{}).items():
    print(k, {a:v.get(a) for a in ("value","ms_per_step","dp_check")}, "e2e", (v.get("e2e") or {}).get("value"), v.get("allreduce"))
PY
