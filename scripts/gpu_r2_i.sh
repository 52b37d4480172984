#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_gpu.py tests/test_full_size_gpu.py tests/test_finetune_gpu.py tests/test_data_gpu.py -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_i.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs"
for CW in direct im2col; do
  VC_CONV1=$CW timeout 300 python bench.py $Q > gpurun_out/i_cfg2_cw$CW.json 2> gpurun_out/i_cfg2_cw$CW.err; echo "cfg2 window=$CW rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/i_cfg2_cw$CW.json'));f=d['families'];print('ms/step %.3f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']), {k:round(f[k]['ms_per_step'],3) for k in ('conv1_1','im2col_rgb','pad_rgb8','conv1_2') if k in f})")"
done
