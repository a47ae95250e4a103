#!/bin/bash
# GPU box: compute-sanitizer over the small mixed workload (fixed + ragged messages, every operation kind)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
