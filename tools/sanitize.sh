#!/usr/bin/env bash
# compute-sanitizer over a small-shape run of every kernel path (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards; synccheck: barrier misuse).  Needs a GPU:  gpurun -- 'bash tools/sanitize.sh'
# Output: gpurun_out/sanitize_<tool>.log  (summaries are copied to profiles/ by hand)
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
for tool in memcheck synccheck racecheck; do
  for part in filter prepass stream; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > "$OUT/sanitize_${tool}_${part}.log" 2>&1
    echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all paths agree' "$OUT/sanitize_${tool}_${part}.log" | tr '\n' ' ')"
  done
done
