#!/bin/bash
# compute-sanitizer over the tcgen05 kernels (scripts/sanitize_mlp.py); logs -> gpurun_out/sanitizer/
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/sanitizer
mkdir -p "$O"
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_mlp.py 700 > "$O/${tool}.log" 2>&1
  echo "$tool rc=$?" | tee -a "$O/summary.txt"
  tail -4 "$O/${tool}.log"
done
ESR_MLP_TILE_OVERLAP=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_mlp.py 20485 > "$O/racecheck_overlap_multitile.log" 2>&1
echo "racecheck_overlap_multitile rc=$?" | tee -a "$O/summary.txt"
tail -4 "$O/racecheck_overlap_multitile.log"
