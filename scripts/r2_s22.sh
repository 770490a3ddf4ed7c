#!/bin/bash
# round 2, session 22: tiny-call parity test, sanitizers after the face-kernel fix (smoke + small test files)
TAG=r2s22
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_tiny_calls.py tests/test_gpu_envelope.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
for tool in memcheck synccheck racecheck; do
  (time timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py --smoke) > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|^real" gpurun_out/${TAG}_sanitizer_$tool.log | tail -4
done
for tool in memcheck racecheck; do
  (time timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_tiny_calls.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py -m gpu -q -x) > gpurun_out/${TAG}_sanitizer_tests_$tool.log 2>&1
  echo "tests $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|^real" gpurun_out/${TAG}_sanitizer_tests_$tool.log | tail -4
done
