#!/bin/bash
# CTA-pair forms of the halo convolution kernels: parity, then the A/B on the bench workloads.
set -u
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_vgg_gpu.py tests/test_gemm_gpu.py tests/test_conv_bwd_gpu.py tests/test_finetune_gpu.py -x -q -m gpu 2>&1 | tail -8
for p in 0 auto 0 auto; do
  if [ $p = auto ]; then unset VC_PAIR; else export VC_PAIR=$p; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pair_$p.json 2> gpurun_out/bench_pair_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pair_$p.json").read().strip().splitlines()[-1])
print("pair=$p", "cfg2 %.3f ms"%d["ms_per_step"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d.get("configs",{}).items()}, d["clocks"])
f=d["families"]
print("   ", {k:round(f[k]["ms_per_step"],3) for k in f if k.startswith("conv") or k.startswith("logits") or k in ("fc","lstm_wgrad","lstm_dx")})
PY
done
