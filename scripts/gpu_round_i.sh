#!/bin/bash
# vocab-projection gradient tiling (split-K wgrad, n-fastest dgrad): train-step parity, then the bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/pytest_train.log 2>&1; echo "train rc=$?"; tail -5 gpurun_out/pytest_train.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_i.json; tail -3 gpurun_out/bench_i.err
