#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --no-cpu-baseline --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
