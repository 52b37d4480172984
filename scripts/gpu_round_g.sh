#!/bin/bash
# Host-side "next" rows on the device (main.py driver round trip with TF bundle checkpoints, Data feature extraction,
# gen_caption), the double-buffered feed, then the default bench and a DRAM-traffic pass over one step.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_main_gpu.py tests/test_data_gpu.py "tests/test_train_step_gpu.py::test_double_buffered_feed_equals_plain_steps" -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_next.log
timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_staged.json; tail -3 gpurun_out/bench_staged.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/traffic_cfg2.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/traffic_cfg2.log 2>&1; echo "ncu traffic rc=$?"; wc -l gpurun_out/traffic_cfg2.csv
