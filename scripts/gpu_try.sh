#!/bin/bash
# scratch: full parity + benches after a kernel change
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
if [ $rc -eq 0 ]; then
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_try_cfg2.json 2> gpurun_out/bench_try_cfg2.err; echo "cfg2 rc=$?"; cut -c1-200 gpurun_out/bench_try_cfg2.json
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-e2e --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_try_cfg4.json 2> gpurun_out/bench_try_cfg4.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_try_cfg4.json
fi
