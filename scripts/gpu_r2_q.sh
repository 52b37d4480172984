#!/bin/bash
# Wide-N (128 output channels, tap halves) form of the halo filter-gradient kernel + the widened CTA-pair policy:
# parity, then cfg 4 A/B (VC_WGRAD_WIDE=0 vs default), then the default bench.
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_conv_bwd_gpu.py tests/test_finetune_gpu.py tests/test_gemm_gpu.py tests/test_vgg_gpu.py tests/test_decode_gpu.py -x -q -m gpu 2>&1 | tail -3
for p in 0 1; do
  VC_WGRAD_WIDE=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs --workload cfg4_finetune_ag_cv_b256 > gpurun_out/bench_cfg4_wide_$p.json 2>gpurun_out/bench_cfg4_wide_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg4_wide_$p.json").read().strip().splitlines()[-1]); f=d["families"]
print("wide=$p", round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in f.items() if k.startswith("wgrad") or k in ("conv_wgrad","conv_dgrad","conv")})
PY
done
bash scripts/gpu_quick.sh "tests/test_abi_cpu.py" "conv1_1,conv1_2,conv2_1,conv2_2,conv3_2,logits_fwd,logits_dgrad,logits_wgrad,ce"
