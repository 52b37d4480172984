#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_dp_gpu.py -m gpu -x -q > gpurun_out/pytest_lstm.log 2>&1; echo "train tests rc=$?"; tail -12 gpurun_out/pytest_lstm.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_train_step_gpu.py --deselect tests/test_dp_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 300 python bench.py --workload cfg3_feats_gmm_cv_b128 --no-cpu-baseline > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; echo "cfg3 rc=$?"; cut -c1-200 gpurun_out/bench_cfg3_n1.json; tail -3 gpurun_out/bench_cfg3_n1.err
