#!/bin/bash
# compute-sanitizer memcheck + racecheck over the round-2 kernels (run on the GPU box): tools/sanitize_r2.sh > gpurun_out/sanitize_r2.log
set -u
for tool in memcheck racecheck; do
  echo "== $tool: round-2 kernels (tools/sanitize_r2.py)"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_r2.py 2>&1 | grep -E "workload ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|assert" | tail -8
done
