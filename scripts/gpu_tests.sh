#!/bin/bash
# usage: gpu_tests.sh [pytest args]  -- GPU parity tests only
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
