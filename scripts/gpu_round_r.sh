#!/bin/bash
# direct conv1_1 (no im2col round trip): VGG parity, then the bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vgg_gpu.py -m gpu -x -q > gpurun_out/pytest_c1.log 2>&1; rc=$?; echo "vgg parity rc=$rc"; tail -3 gpurun_out/pytest_c1.log
if [ $rc -eq 0 ]; then
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
fi
