#!/bin/bash
# compute-sanitizer over the hot path (SURVEY 5 "race detection"): memcheck (out-of-bounds / misaligned global, shared,
# TMA-written memory) and racecheck (shared-memory hazards between the warp roles of the persistent kernels: TMA producer,
# MMA issuer, I/O warp, epilogue warps) on small train steps of every prior and one greedy + beam decode.
# Logs: gpurun_out/sanitize_{memcheck,racecheck}.log -- copy the summaries to profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for TOOL in memcheck racecheck; do
  timeout 1500 $CS --tool $TOOL --error-exitcode 9 --print-limit 20 python scripts/sanitize_case.py > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|train step ok|kernel level ok|decode ok|Error|error" gpurun_out/sanitize_$TOOL.log | head -12
done
