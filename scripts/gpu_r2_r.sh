#!/bin/bash
# ReLU derivative + bias gradient fused into the input-gradient GEMM epilogues: parity, cfg 4 A/B (VC_FUSE_RELU_BWD=0 vs default)
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_finetune_gpu.py tests/test_conv_bwd_gpu.py tests/test_vgg_gpu.py -x -q -m gpu 2>&1 | tail -3
for p in 0 1; do
  VC_FUSE_RELU_BWD=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs --workload cfg4_finetune_ag_cv_b256 > gpurun_out/bench_cfg4_fuse_$p.json 2>gpurun_out/bench_cfg4_fuse_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg4_fuse_$p.json").read().strip().splitlines()[-1]); f=d["families"]
print("fuse=$p", round(d["ms_per_step"],3), {k:(round(v["ms_per_step"],3), v["launches_per_step"]) for k,v in f.items() if k in ("relu_bwd","pool_relu_bwd","conv_dgrad","conv_wgrad","conv")}, {k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("dgrad")})
PY
done
