#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
bash scripts/gpu_scale.sh $N
VC_GRAD_BF16=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --workload cfg3_feats_gmm_cv_b128 --no-cpu-baseline --no-profile > gpurun_out/bench_cfg3_n${N}_bf16.json 2> gpurun_out/bench_cfg3_n${N}_bf16.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg3_n${N}_bf16.json").read()); a=d["allreduce"]
print("cfg3 bf16 transport: ms/step %.3f value %.0f no_ar %.3f exposed %.3f"%(d["ms_per_step"],d["value"],a["ms_per_step_no_allreduce"],a["exposed_ms"]), d["dp_check"], a["transport"])
PY
