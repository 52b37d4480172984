#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_gpu.py tests/test_data_gpu.py "tests/test_full_size_gpu.py::test_vgg_forward_rows_independent_and_sampled_rows_match_oracle" -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_j.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs"
for CW in direct im2col; do
  VC_CONV1=$CW timeout 300 python bench.py $Q > gpurun_out/j_cfg2_$CW.json 2> gpurun_out/j_cfg2_$CW.err; echo "cfg2 conv1=$CW rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/j_cfg2_$CW.json'));f=d['families'];print('ms/step %.3f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']), {k:round(f[k]['ms_per_step'],3) for k in ('conv1_1','im2col_rgb','conv1_2') if k in f})")"
done
