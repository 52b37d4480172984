#!/bin/bash
# Full GPU suite, the default bench exactly as the driver calls it (CPU baseline leg on), the reference arm,
# and the ncu launch list of the same command.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
