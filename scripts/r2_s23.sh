#!/bin/bash
# round 2, session 23: kd facet order (surface_order=2) against Hilbert: parity tests, envelope / nearest / faces A/B; tiny-call tests under sanitizers
TAG=r2s23
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_tiny_calls.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_tiny.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_tiny.log
(time TWG_SURFACE_ORDER=2 timeout 900 python -m pytest tests/test_gpu_envelope.py tests/test_gpu_robustness.py tests/test_gpu_tiny_calls.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_kd.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_kd.log
for o in 1 2; do
  TWG_SURFACE_ORDER=$o timeout 900 python bench.py --parts envelope,nearest,envelope_faces,envelope_faces_c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_order$o.log 2>&1
done
for tool in memcheck racecheck; do
  (time timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_tiny_calls.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py -m gpu -q -x) > gpurun_out/${TAG}_sanitizer_tests_$tool.log 2>&1
  echo "tests $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|^real" gpurun_out/${TAG}_sanitizer_tests_$tool.log | tail -4
done
python - <<'PY'
import json
for o in (1, 2):
    for l in open('gpurun_out/r2s23_order%d.log' % o):
        if l.startswith('{'):
            d = json.loads(l)
            print('order', o, 'envelope %.3f ms' % d['ms_per_step'], 'mism', d['extra'].get('decision_mismatches_vs_oracle_100k_sample'), 'build', d['extra'].get('surface_build_s'))
            for k, p in d['parts'].items(): print('   ', k, '%.3f ms %.3e' % (p['ms_per_step'], p['value']), {a: b for a, b in p['extra'].items() if 'mism' in a or 'parity' in a or 'near_surface' in a})
PY
