#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_full_size_gpu.py tests/test_dp_gpu.py tests/test_finetune_gpu.py tests/test_main_gpu.py -m gpu -q -x > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_e.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --no-profile"
for OV in 1 0; do
for W in cfg3_feats_gmm_cv_b128 feats_normal_b256 cfg2_vgg_normal_b256 cfg4_finetune_ag_cv_b256; do
  VC_BWD_OVERLAP=$OV timeout 300 python bench.py --workload $W $Q > gpurun_out/e_${W}_ov$OV.json 2> gpurun_out/e_${W}_ov$OV.err; echo "$W ov=$OV rc=$? $(python -c "import json;d=json.load(open('gpurun_out/e_${W}_ov$OV.json'));print('ms/step %.3f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']))")"
done
done
