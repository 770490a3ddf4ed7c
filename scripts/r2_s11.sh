#!/bin/bash
# round 2, session 11: device-side winding hierarchy build: bit-identity tests, build time, one-shot bench
TAG=r2s11
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_winding.py tests/test_truth.py tests/test_gpu_multi.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_winding.log 2>&1
tail -12 gpurun_out/${TAG}_pytest_winding.log
timeout 600 python - > gpurun_out/${TAG}_build_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, tetwild_b200 as tw
from tetwild_b200 import synth
for name, (V, F) in (("sphere 1.0M", synth.uv_sphere(708, 708)), ("sphere 100k", synth.uv_sphere(224, 224)), ("icosphere 20k", synth.icosphere(5))):
    for dev in (1, 0):
        c = tw.Context(0); c.set_option("winding_device_build", dev)
        ts = []
        for _ in range(4):
            t = time.perf_counter(); W = tw.Winding(c, V, F); c.synchronize(); ts.append(time.perf_counter() - t); st = W.stats(); W.close()
        print("%-14s device_build=%d: %.1f ms (first %.1f ms) %s" % (name, dev, min(ts[1:]) * 1e3, ts[0] * 1e3, st))
        c.close()
PY
cat gpurun_out/${TAG}_build_time.log
timeout 900 python bench.py --parts winding_oneshot,winding --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_oneshot.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s11_oneshot.log'):
    if l.startswith('{'):
        d=json.loads(l); print('oneshot', '%.3e q/s'%d['value'], '%.1f ms'%d['ms_per_step'], d['extra'])
        for k,p in d['parts'].items(): print(k,'%.3e'%p['value'], p['extra'].get('hierarchy_build_s'), p['roofline'].get('pairs_per_query'))
PY
