#!/bin/bash
# round 2, session 16: lane-per-ring kernel: size sweep, wide vertex loads, ncu at full size
TAG=r2s16
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_amips.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for sc in 0.1 0.32 0.64 1.0; do
  for m in 0 1 2; do
    TWG_RING_MODE=$m timeout 600 python bench.py --parts amips_ring --steps 6 --warmup 3 --no-cpu --scale $sc > gpurun_out/${TAG}_ring_m${m}_s$sc.log 2>&1
  done
done
TWG_RING_MODE=1 timeout 600 ncu --set full --metrics l1tex__t_bytes.sum,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:amips_ring_lane -c 1 -f -o gpurun_out/${TAG}_ring_lane50 python scripts/prof_part.py ring 50000000 2 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py rep gpurun_out/${TAG}_ring_lane50.ncu-rep gpurun_out/${TAG}_ring_lane50.txt
head -40 gpurun_out/${TAG}_ring_lane50.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s16_ring_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('/')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'])
PY
