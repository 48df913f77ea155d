#!/bin/bash
# compute-sanitizer memcheck + racecheck over small shapes of every kernel family (run on the GPU box).
# usage: tools/sanitize.sh > gpurun_out/sanitize.log
set -u
K='forward_and_inverse_vs_oracle and (3] or 7] or 10] or 11] or 12])'
for tool in memcheck racecheck; do
  echo "== $tool: 768-bit"; timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_ntt768.py -q -x -k "$K or batched_and_strided" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -4
  echo "== $tool: 32-bit"; timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_ntt32.py -q -x -k "forward_and_inverse_vs_oracle and (5] or 12] or 16] or 17] or 18] or 20])" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -4
  echo "== $tool: widened + fused four-step"; timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_widened.py tests/test_gpu_fourstep.py -q -x -k "inner_product and (129 or 5000) or powers or coset and 10 or fused and 12" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -4
done
