#!/bin/bash
# Round-end validation exactly as the driver runs it: full GPU suite, smoke, default bench (CPU baseline leg on), reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
