#!/bin/bash
# 2-rank run: DP tests + the bench under torchrun (device leg pipelined, e2e leg through the staged feed + all-reduce)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py "tests/test_train_step_gpu.py::test_double_buffered_feed_equals_plain_steps" -m gpu -x -q > gpurun_out/pytest_dp.log 2>&1; echo "dp rc=$?"; tail -4 gpurun_out/pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; echo "n2 rc=$?"; cut -c1-300 gpurun_out/bench_cfg2_n2.json; tail -3 gpurun_out/bench_cfg2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg3_feats_gmm_cv_b128 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_cfg3_n2.err; echo "cfg3 n2 rc=$?"; cut -c1-300 gpurun_out/bench_cfg3_n2.json; tail -3 gpurun_out/bench_cfg3_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json
