#!/bin/bash
# round 2, session 20: wide vertex loads in the quality pass (A/B), sort key width 27 / 30, nearest near-surface subset, one-shot winding
TAG=r2s20
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
TWG_WIDE_GATHER=1 timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -q -x >> gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for w in 0 1; do
  TWG_WIDE_GATHER=$w timeout 600 python bench.py --parts amips_quality --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_quality_w$w.log 2>&1
done
for b in 27 30; do
  TWG_SORT_BITS=$b timeout 600 python bench.py --parts envelope --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_sortbits$b.log 2>&1
done
timeout 900 python bench.py --parts nearest,winding_oneshot --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_near_oneshot.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s20_quality*.log')) + sorted(glob.glob('gpurun_out/r2s20_sortbits*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('/')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['roofline'].get('hbm_frac'))
for l in open('gpurun_out/r2s20_near_oneshot.log'):
    if l.startswith('{'):
        d = json.loads(l); print('nearest %.3e' % d['value'], d['extra'].get('near_surface_subset'))
        p = d['parts']['winding_oneshot']; print('oneshot %.3e %.1f ms' % (p['value'], p['ms_per_step']), p['extra'])
PY
