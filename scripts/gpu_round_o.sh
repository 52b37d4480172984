#!/bin/bash
set -u
mkdir -p gpurun_out
VC_WGRAD_HALO_MAXCH=256 timeout 300 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -x -q > gpurun_out/pytest_wg256.log 2>&1; echo "wgrad256 parity rc=$?"; tail -3 gpurun_out/pytest_wg256.log
VC_WGRAD_HALO_MAXCH=256 timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-e2e --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_cfg4_wg256.json 2> gpurun_out/bench_cfg4_wg256.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_cfg4_wg256.json; tail -3 gpurun_out/bench_cfg4_wg256.err
