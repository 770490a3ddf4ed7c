#!/bin/bash
# round 2, session 17: lane kernel with contiguous ring ranges; one-launch tiny calls (parameters + in-kernel completion word)
TAG=r2s17
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py tests/test_gpu_smoothing_pass.py tests/test_gpu_envelope.py tests/test_gpu_robustness.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/latency.py > gpurun_out/${TAG}_latency.log 2>&1; cat gpurun_out/${TAG}_latency.log | cut -c1-700
for sc in 0.32 1.0; do
  for m in 1 2; do
    TWG_RING_MODE=$m timeout 600 python bench.py --parts amips_ring --steps 6 --warmup 3 --no-cpu --scale $sc > gpurun_out/${TAG}_ring_m${m}_s$sc.log 2>&1
  done
done
timeout 600 python bench.py --parts pass_stream --steps 3 --warmup 1 --no-cpu > gpurun_out/${TAG}_pass.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s17_ring_*.log')) + ['gpurun_out/r2s17_pass.log']:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('/')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], json.dumps(d.get('extra', {}))[:600] if 'pass' in f else '')
PY
