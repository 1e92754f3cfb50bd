#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the smoke invocation of every kernel; small shapes only.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
