#!/bin/bash
# round 2, session 18: one-launch tiny calls incl. faces; latency table with the empty-kernel floor; ring kernel restored
TAG=r2s18
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py tests/test_gpu_smoothing_pass.py tests/test_gpu_envelope.py tests/test_gpu_robustness.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/latency.py > gpurun_out/${TAG}_latency.log 2>&1; cat gpurun_out/${TAG}_latency.log | cut -c1-900
timeout 600 python bench.py --parts amips_ring,pass_stream --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_pass.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s18_ring_pass.log'):
    if l.startswith('{'):
        d = json.loads(l); print('ring %.3f ms %.3e' % (d['ms_per_step'], d['value']))
        p = d['parts']['pass_stream']; print('pass', '%.3e' % p['value'], json.dumps(p.get('extra', {}))[:1500])
PY
