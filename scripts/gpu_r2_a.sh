#!/bin/bash
# Round 2, first GPU pass: whole GPU suite, default bench (with the cfg 3 / cfg 4 side workloads), LSTM per-step timelines.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
Q="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-profile --no-extra-configs"
VC_LSTM_TRACE=1 timeout 200 python bench.py --workload feats_normal_b256 $Q > gpurun_out/trace_n1280.json 2> gpurun_out/trace_n1280.err; echo "trace rc=$?"
VC_LSTM_TRACE=1 timeout 200 python bench.py --workload cfg3_feats_gmm_cv_b128 $Q > gpurun_out/trace_n640.json 2> gpurun_out/trace_n640.err; echo "trace rc=$?"
grep -c "lstm trace" gpurun_out/trace_n1280.err gpurun_out/trace_n640.err
