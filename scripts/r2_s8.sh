#!/bin/bash
# round 2, session 8: nearest (packet, Hilbert facets), query sort curve A/B, ring kernel at 64 registers, call stream, new gpu tests
TAG=r2s8
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[2]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['metric'][:28], '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], {k:v for k,v in d['extra'].items() if 'mism' in k or 'parity' in k})
        for k,p in d.get('parts',{}).items(): print('   ',k,'%.3e'%p['value'],'%.3f ms'%p['ms_per_step'], 'e2e %.3e'%p['e2e']['value'], {a:b for a,b in p['extra'].items() if 'mism' in a or 'parity' in a})
PY
}
(time timeout 1200 python -m pytest tests/test_gpu_callstream.py tests/test_gpu_amips.py tests/test_gpu_envelope.py tests/test_truth.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for C in 0 1; do
TWG_SORT_CURVE=$C timeout 600 python bench.py --parts envelope,nearest --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_curve$C.log 2>&1
show "sort_curve=$C" gpurun_out/${TAG}_curve$C.log
done
TWG_NEAREST_MODE=2 timeout 600 python bench.py --parts nearest --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_near_rounds.log 2>&1
show "nearest rounds" gpurun_out/${TAG}_near_rounds.log
TWG_RING_MINB=4 timeout 600 python bench.py --parts amips_ring --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_minb4.log 2>&1
show "ring minb=4" gpurun_out/${TAG}_ring_minb4.log
timeout 600 python bench.py --parts pass_stream --steps 3 --warmup 1 > gpurun_out/${TAG}_stream.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s8_stream.log'):
    if l.startswith('{'):
        d=json.loads(l); print('pass_stream', '%.3e calls/s batched'%d['value'], d['extra'], 'cpu', d['cpu_baseline']['value'])
PY
python scripts/latency.py > gpurun_out/${TAG}_latency.log 2>&1; tail -12 gpurun_out/${TAG}_latency.log
