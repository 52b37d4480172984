#!/bin/bash
# 2-GPU call: DP test (1 GPU), cfg5 decode bench, then the 2-rank bench (cfg2) and cfg3 at 1 and 2 ranks
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_decode_gpu.py -m gpu -x -q > gpurun_out/pytest_dp.log 2>&1; echo "dp/decode tests rc=$?"; tail -4 gpurun_out/pytest_dp.log
timeout 900 python bench.py --workload cfg5_decode_greedy_beam5 --steps 3 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-700 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
RUN2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN2 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; echo "cfg2 n2 rc=$?"; cut -c1-300 gpurun_out/bench_cfg2_n2.json; tail -3 gpurun_out/bench_cfg2_n2.err
timeout 600 python bench.py --workload cfg3_feats_gmm_cv_b128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; echo "cfg3 n1 rc=$?"; cut -c1-300 gpurun_out/bench_cfg3_n1.json; tail -3 gpurun_out/bench_cfg3_n1.err
timeout 300 $RUN2 bench.py --gpus 2 --workload cfg3_feats_gmm_cv_b128 --steps 20 --warmup 5 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err; echo "cfg3 n2 rc=$?"; cut -c1-300 gpurun_out/bench_cfg3_n2.json; tail -3 gpurun_out/bench_cfg3_n2.err
