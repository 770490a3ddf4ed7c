#!/bin/bash
# round 2, session 21: compute-sanitizer over the smoke invocation (every entry point at small sizes): memcheck, racecheck, synccheck
TAG=r2s21
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  (time timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py --smoke) > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|^real" gpurun_out/${TAG}_sanitizer_$tool.log | tail -4
done
