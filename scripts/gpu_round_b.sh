#!/bin/bash
# fine-tune parity tests first (new kernels, short timeout), then the whole GPU suite, then benches
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_finetune_gpu.py -m gpu -x -q > gpurun_out/pytest_ft.log 2>&1; echo "finetune rc=$?"; tail -25 gpurun_out/pytest_ft.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_finetune_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e --workload finetune_ag_cv_b64 --steps 5 > gpurun_out/bench_ft64.json 2> gpurun_out/bench_ft64.err; echo "bench ft64 rc=$?"; cut -c1-300 gpurun_out/bench_ft64.json; tail -3 gpurun_out/bench_ft64.err
timeout 600 python bench.py --no-cpu-baseline --workload cfg4_finetune_ag_cv_b256 --steps 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; cut -c1-300 gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
