#!/bin/bash
# ncu evidence for the current code (one GPU, never under torchrun):
#   1. launch list of the bench command (per-launch gpu__time_duration; kernel shares of the step)      -> launches_cfg2.csv
#   2. DRAM bytes per launch for one step (feeds scripts/traffic_from_ncu.py -> profiles/traffic.json)    -> traffic_cfg2.csv
#   3. --set full captures of the kernels named on the command line, default: the two halo convolutions and the
#      halo-form filter gradient (cfg 4)
# usage: gpurun -- 'bash scripts/gpu_profiles.sh [name:regex:skip[:workload] ...]'; copy what matters into profiles/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-profile"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv \
   $BENCH --steps 2 > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
   --log-file gpurun_out/traffic_cfg2.csv $BENCH --steps 1 > gpurun_out/traffic_cfg2.log 2>&1; echo "ncu traffic rc=$?"; wc -l gpurun_out/traffic_cfg2.csv
if [ $# -eq 0 ]; then
  set -- conv2_2_halo_stream:conv_halo_stream_kernel:3 conv1_2_halo:conv_halo_kernel:6 wgrad1_2_halo:conv_wgrad_halo_kernel:8:cfg4_finetune_ag_cv_b256
fi
bash scripts/gpu_ncu.sh "$@"
