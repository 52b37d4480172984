#!/bin/bash
set -u
mkdir -p gpurun_out
VC_VOCAB_CHUNK=256 timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_dp_gpu.py -m gpu -q -x > gpurun_out/pytest_k1.log 2>&1; echo "pytest(chunk 256) rc=$?"; tail -2 gpurun_out/pytest_k1.log
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_comm_gpu.py -m gpu -q -x > gpurun_out/pytest_k2.log 2>&1; echo "pytest(full size) rc=$?"; tail -2 gpurun_out/pytest_k2.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --no-profile --no-e2e"
for CH in 2048 0 1024 4096; do
for W in feats_normal_b256 cfg3_feats_gmm_cv_b128; do
  VC_VOCAB_CHUNK=$CH timeout 300 python bench.py --workload $W $Q > gpurun_out/k_${W}_$CH.json 2> gpurun_out/k_${W}_$CH.err; echo "$W chunk=$CH rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/k_${W}_$CH.json'));print('ms/step %.3f value %.0f'%(d['ms_per_step'],d['value']))")"
done
done
