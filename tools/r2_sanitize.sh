#!/bin/bash
# compute-sanitizer on the final kernels: memcheck + racecheck + synccheck on BASELINE.json configs[0]
# (5k / 256^2, init opacity) and on the big-tile case, forward + backward through the public API.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|cases ok" gpurun_out/r2_sanitizer_$tool.log | tail -4
done
