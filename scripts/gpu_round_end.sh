#!/bin/bash
# Round-end evidence next to the validation script: the other workloads' bench lines, the ncu launch list of the bench
# command, and --set full captures of the vocab filter gradient and of one step's HBM-bound kernels.
set -u
bash scripts/gpu_validate.sh
B="timeout 100 python bench.py --gpus 1 --no-cpu-baseline"
$B --workload cfg3_feats_gmm_cv_b128 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; echo "cfg3 rc=$?"; cut -c1-180 gpurun_out/bench_cfg3_n1.json
$B --workload cfg4_finetune_ag_cv_b256 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-180 gpurun_out/bench_cfg4.json
BENCH="python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-profile"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv $BENCH --steps 2 > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
timeout 200 bash scripts/gpu_ncu.sh fc1:EpiStore:66   # the 67th EpiStore GEMM of the run is fc1 of the 4th step (grid 128, 205 MB of weights)
NCU_COUNT=12 timeout 200 bash scripts/gpu_ncu.sh hbm_kernels:'k_ce|k_sample_z|k_dz_reduce|k_adam|k_sumsq|k_embed_scatter|k_colsum|k_kl_rows':36
