#!/bin/bash
# ncu evidence for the current code: launch list of the bench command + one --set full capture each of the streamed-filter
# halo convolution (conv2_2) and of the resident-filter halo convolution (conv1_2)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/launches_cfg2.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_cfg2.csv
bash scripts/gpu_ncu.sh conv2_2_halo_stream:conv_halo_stream_kernel:3 conv1_2_halo:conv_halo_kernel:6
