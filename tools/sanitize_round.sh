#!/bin/bash
# compute-sanitizer over every kernel and variant on a tiny case (tools/sanitize_case.py), all four tools; logs under gpurun_out/,
# copied into profiles/ when clean.  usage (on the GPU box): bash tools/sanitize_round.sh r2
P=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_case.py > gpurun_out/${P}_sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${P}_sanitize_rc.log
done
