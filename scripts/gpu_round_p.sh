#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_decode_gpu.py -m gpu -x -q --durations=5 > gpurun_out/pytest_full.log 2>&1; echo "full-size + decode rc=$?"; tail -14 gpurun_out/pytest_full.log
timeout 300 python scripts/debug/beam_idem.py 4.0 2>&1 | tail -6
timeout 300 python bench.py --workload cfg5_decode_greedy_beam5 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-250 gpurun_out/bench_cfg5.json; tail -2 gpurun_out/bench_cfg5.err
