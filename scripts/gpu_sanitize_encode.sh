#!/bin/bash
# compute-sanitizer over the run-merged encode backward (warp shuffles under full masks, per-thread shared-memory columns):
# the golden-size A/B test against the plain scatter under memcheck / synccheck / racecheck; logs -> gpurun_out/sanitizer_encode/
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/sanitizer_encode
mkdir -p "$O"
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_voxurff.py -m gpu -q -x \
      -k "merged_reds and golden" > "$O/${tool}.log" 2>&1
  echo "$tool rc=$?" | tee -a "$O/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$O/${tool}.log" | tail -3
done
