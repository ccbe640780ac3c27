#!/bin/bash
# tools/sanitize.sh -- compute-sanitizer memcheck / racecheck / synccheck over the smoke configuration and a multi-slice CABAC batch.
# Run on the GPU box: `gpurun -- bash tools/sanitize.sh r02`; logs land in gpurun_out/, summaries are copied to profiles/ by hand.
tag=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
python tools/sanitize_case.py all > gpurun_out/sanitize_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain_$tag.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  for case in smoke cabac; do
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_case.py $case > gpurun_out/sanitize_${tool}_${case}_$tag.log 2>&1
    echo "$tool $case exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${case}_$tag.log | tail -1)"
  done
done
