#!/bin/bash
# compute-sanitizer (memcheck + racecheck; initcheck and synccheck are run the same way) over scripts/sanitize_job.py: every kernel and launch path at small shapes.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_job.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -5 gpurun_out/sanitizer_$tool.log
done
