#!/bin/bash
# Frozen-extractor pipelining: VGG16 forward of batch i+1 on the copy / side stream under the caption model of batch i
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_train_step_gpu.py::test_double_buffered_feed_equals_plain_steps" tests/test_main_gpu.py tests/test_vgg_gpu.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pipe rc=$?"; tail -5 gpurun_out/pytest_pipe.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_j.json; tail -3 gpurun_out/bench_j.err
