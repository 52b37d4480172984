#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_gpu.py tests/test_finetune_gpu.py tests/test_conv_bwd_gpu.py tests/test_data_gpu.py -m gpu -q -x > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_l.log
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs --no-e2e"
for W in cfg2_vgg_normal_b256 cfg4_finetune_ag_cv_b256; do
  timeout 300 python bench.py --workload $W $Q > gpurun_out/l_$W.json 2> gpurun_out/l_$W.err; echo "$W rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/l_$W.json'));f=d['families'];print('ms/step %.3f value %.0f'%(d['ms_per_step'],d['value']), {k:(round(f[k]['ms_per_step'],3), f[k]['launches_per_step']) for k in ('conv1_1','im2col_rgb','refresh_shadows') if k in f})")"
done
