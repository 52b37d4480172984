#!/bin/bash
# scratch: LSTM parity, timeline trace, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_dp_gpu.py tests/test_main_gpu.py -m gpu -x -q > gpurun_out/pytest_lstm.log 2>&1; rc=$?; echo "lstm parity rc=$rc"; tail -3 gpurun_out/pytest_lstm.log
if [ $rc -eq 0 ]; then
  VC_LSTM_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile 2>&1 | grep -A8 "lstm trace.*#[45]" | head -40
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_try_cfg2.json 2> gpurun_out/bench_try_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 gpurun_out/bench_try_cfg2.json
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-e2e --workload cfg3_feats_gmm_cv_b128 > gpurun_out/bench_try_cfg3.json 2> gpurun_out/bench_try_cfg3.err; echo "cfg3 rc=$?"; cut -c1-200 gpurun_out/bench_try_cfg3.json
fi
