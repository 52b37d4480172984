#!/bin/bash
# Halo-form filter gradient: per-layer parity vs torch autograd, fine-tune step parity, cfg4 bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -x -q > gpurun_out/pytest_wg.log 2>&1; rc=$?; echo "wgrad parity rc=$rc"; tail -6 gpurun_out/pytest_wg.log
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_finetune_gpu.py -m gpu -x -q > gpurun_out/pytest_ft.log 2>&1; echo "finetune rc=$?"; tail -4 gpurun_out/pytest_ft.log
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_cfg4_wg.json 2> gpurun_out/bench_cfg4_wg.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_cfg4_wg.json; tail -3 gpurun_out/bench_cfg4_wg.err
fi
