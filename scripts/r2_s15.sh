#!/bin/bash
# round 2, session 15: lane-per-ring one-ring kernel: parity, A/B, ncu
TAG=r2s15
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py tests/test_gpu_smoothing_pass.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
for m in 0 1; do
  TWG_RING_MODE=$m timeout 600 python bench.py --parts amips_ring --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_mode$m.log 2>&1
done
for w in 2 4 6; do
  TWG_RING_MODE=1 TWG_RING_WAVES=$w timeout 600 python bench.py --parts amips_ring --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_mode1_w$w.log 2>&1
done
TWG_RING_MODE=1 timeout 600 ncu --set full --metrics l1tex__t_bytes.sum,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:amips_ring_lane -c 1 -f -o gpurun_out/${TAG}_ring_lane python scripts/prof_part.py ring 16000000 2 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py rep gpurun_out/${TAG}_ring_lane.ncu-rep gpurun_out/${TAG}_ring_lane.txt
head -40 gpurun_out/${TAG}_ring_lane.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s15_ring_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('/')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['roofline'].get('hbm_frac'))
PY
