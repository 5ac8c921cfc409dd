#!/bin/bash
# compute-sanitizer over tools/sanitizer_cases.py: memcheck, racecheck, synccheck.  Logs land in
# gpurun_out/sanitizer_<tool>_<group>.log; the last lines of each carry the tool's error summary.
#   usage: tools/run_sanitizer.sh [timeout-seconds-per-run]
T=${1:-600}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for group in attn gemm engine; do
    log=gpurun_out/sanitizer_${tool}_${group}.log
    timeout $T compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_cases.py $group > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool $group: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $log | tr '\n' ' ')"
  done
done
