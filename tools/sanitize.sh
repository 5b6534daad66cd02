#!/bin/bash
# usage (GPU box): tools/sanitize.sh [tiles]  ->  gpurun_out/sanitize_{memcheck,racecheck}.log
# compute-sanitizer on hist_kernel and encode_kernel (packed, packed + check, wide; multi-tile inputs).
tiles=${1:-24}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $tiles > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.log
done
