#!/bin/bash
# round 2, session 7: one-ring prefetch pipeline A/B; Hilbert vs Morton facet order (envelope, nearest, faces); gpu tests
TAG=r2s7
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[2]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['metric'][:28], '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], {k:v for k,v in d['extra'].items() if 'mism' in k or 'parity' in k})
        for k,p in d.get('parts',{}).items(): print('   ',k,'%.3e'%p['value'],'%.3f ms'%p['ms_per_step'], {a:b for a,b in p['extra'].items() if 'mism' in a or 'parity' in a})
PY
}
for A in 1 0; do
TWG_RING_PREFETCH=$A timeout 600 python bench.py --parts amips_ring --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_pf$A.log 2>&1
show "ring prefetch=$A" gpurun_out/${TAG}_ring_pf$A.log
done
for H in 1 0; do
TWG_SURFACE_ORDER=$H timeout 600 python bench.py --parts envelope,nearest,envelope_faces --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_order$H.log 2>&1
show "surface_order=$H" gpurun_out/${TAG}_order$H.log
done
(time timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:amips_ring_pf -c 1 -o gpurun_out/${TAG}_ring python bench.py --parts amips_ring --steps 1 --warmup 3 --no-cpu --scale 0.32 > gpurun_out/${TAG}_ncu_ring.log 2>&1
