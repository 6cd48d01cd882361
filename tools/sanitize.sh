#!/bin/bash
# Round sanitizer run (GPU box): racecheck + synccheck + memcheck over every tcgen05 kernel variant.
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/t_sanitize.py > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "$tool rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -3
done
