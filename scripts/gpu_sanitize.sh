#!/bin/bash
# compute-sanitizer over smoke() (tiny config: one forward + a 4-step DDIM sample through the C ABI)
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|error" gpurun_out/sanitizer_$tool.log | head -8
done
