#!/bin/bash
# N-rank sweep of the knobs that decide how much of the gradient all-reduce hides under the backward pass (cfg 3 shapes)
set -u
N=${1:-8}
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --workload cfg3_feats_gmm_cv_b128 --no-cpu-baseline --no-profile > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_$tag.json").read()); a=d["allreduce"]
    print("$tag", "ms/step %.3f"%d["ms_per_step"], "no_ar %.3f unbucketed %.3f exposed %.3f span %.3f"%(a["ms_per_step_no_allreduce"],a["ms_per_step_unbucketed"],a["exposed_ms"],a["span_ms"]), d["dp_check"])
except Exception as e:
    print("$tag ERR", e)
PY
}
run base A=1
run side32 VC_SIDE_SMS=32
run side16 VC_SIDE_SMS=16
run side32_cta16 VC_SIDE_SMS=32 NCCL_MAX_CTAS=16
run side32_cta8 VC_SIDE_SMS=32 NCCL_MAX_CTAS=8
run nooverlap VC_BWD_OVERLAP=0
run nooverlap_cta16 VC_BWD_OVERLAP=0 NCCL_MAX_CTAS=16
