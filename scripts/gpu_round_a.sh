#!/bin/bash
# GPU tests, then the conv "halo" experiment: VGG parity + bench with VC_CONV_HALO=1 (both descriptor modes)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for bo in 1 0; do
  VC_CONV_HALO=1 VC_HALO_BASE_OFFSET=$bo timeout 600 python -m pytest tests/test_vgg_gpu.py -m gpu -x -q > gpurun_out/pytest_halo_$bo.log 2>&1; echo "halo bo=$bo rc=$?"; tail -12 gpurun_out/pytest_halo_$bo.log
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_nohalo.json 2> gpurun_out/bench_nohalo.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_nohalo.json
VC_CONV_HALO=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_halo.json 2> gpurun_out/bench_halo.err; echo "bench halo rc=$?"; cut -c1-400 gpurun_out/bench_halo.json
