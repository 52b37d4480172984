#!/bin/bash
# Streamed-filter halo convolution (conv2_2 + input gradients of conv2_1 / conv2_2 / conv3_1): parity first, then speed
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_bwd_gpu.py tests/test_vgg_gpu.py -m gpu -x -q > gpurun_out/pytest_halo2.log 2>&1; rc=$?; echo "halo2 parity rc=$rc"; tail -6 gpurun_out/pytest_halo2.log
if [ $rc -ne 0 ]; then
  VC_CONV_HALO2=0 timeout 300 python -m pytest tests/test_conv_bwd_gpu.py tests/test_vgg_gpu.py -m gpu -x -q > gpurun_out/pytest_halo2_off.log 2>&1; echo "halo2 OFF parity rc=$?"; tail -4 gpurun_out/pytest_halo2_off.log
fi
timeout 600 python -m pytest tests/test_main_gpu.py tests/test_data_gpu.py tests/test_finetune_gpu.py "tests/test_train_step_gpu.py::test_double_buffered_feed_equals_plain_steps" -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "next rc=$?"; tail -8 gpurun_out/pytest_next.log
if [ $rc -eq 0 ]; then
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_halo2.json 2> gpurun_out/bench_halo2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_halo2.json; tail -3 gpurun_out/bench_halo2.err
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_cfg4_halo2.json 2> gpurun_out/bench_cfg4_halo2.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_cfg4_halo2.json; tail -3 gpurun_out/bench_cfg4_halo2.err
fi
