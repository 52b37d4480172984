#!/bin/bash
# usage: gpu_prof.sh "<name>:<gemm ordinal>" ...   -- ncu --set full of single gemm_tc_kernel launches (by ordinal) of the bench step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile"
for spec in "$@"; do
  IFS=: read name ord <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k gemm_tc_kernel -s $ord -c 1 -f -o gpurun_out/prof_$name $BENCH > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json
ls gpurun_out
