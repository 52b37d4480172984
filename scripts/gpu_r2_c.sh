#!/bin/bash
set -u
mkdir -p gpurun_out
VC_LSTM_UNITS=64 timeout 300 python -m pytest tests/test_train_step_gpu.py -m gpu -q -x -k "forward_normal_prior or gradients" > gpurun_out/pytest_u64.log 2>&1; echo "units64 rc=$?"; tail -3 gpurun_out/pytest_u64.log
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_full_size_gpu.py tests/test_dp_gpu.py tests/test_main_gpu.py -m gpu -q > gpurun_out/pytest_lstm2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_lstm2.log
