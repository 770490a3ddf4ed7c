#!/bin/bash
# round 2, session 26: kd order with the split axis chosen by child surface area (surface_order=3) against the longest-axis kd order (2); new defaults
TAG=r2s26
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_winding.py tests/test_gpu_envelope.py -m gpu -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for o in 2 3; do
  TWG_SURFACE_ORDER=$o timeout 900 python bench.py --parts envelope,nearest,envelope_faces,envelope_faces_c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_order$o.log 2>&1
done
python - <<'PY'
import json
for o in (2, 3):
    for l in open('gpurun_out/r2s26_order%d.log' % o):
        if l.startswith('{'):
            d = json.loads(l)
            print('order', o, 'envelope %.3f ms' % d['ms_per_step'], 'mism', d['extra'].get('decision_mismatches_vs_oracle_100k_sample'))
            for k, p in d['parts'].items(): print('   ', k, '%.3f ms %.3e' % (p['ms_per_step'], p['value']), p['extra'].get('near_surface_subset', {}).get('points_per_s'), p['extra'].get('parity_vs_brute_force', p['extra'].get('decision_mismatches_vs_oracle_sample')))
PY
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, tetwild_b200 as tw
from tetwild_b200 import synth
V, F = synth.torus_knot(); V = synth.normalise_unit_diag(V)
for o in (1, 2, 3):
    c = tw.Context(0); c.set_option("surface_order", o)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); S = tw.Surface(c, V, F); c.synchronize(); ts.append(time.perf_counter() - t); S.close()
    print("surface build, %d facets, surface_order=%d: %.1f ms (first %.1f ms)" % (len(F), o, min(ts[1:]) * 1e3, ts[0] * 1e3))
    c.close()
PY
