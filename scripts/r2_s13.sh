#!/bin/bash
# round 2, session 13: member-stream one-ring kernel (parity + A/B), build phase list, sort-bits A/B
TAG=r2s13
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_callstream.py tests/test_gpu_winding.py tests/test_gpu_smoothing_pass.py -m gpu -q -x) > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log
for m in 0 1 2; do
  TWG_RING_MODE=$m timeout 600 python bench.py --parts amips_ring --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_mode$m.log 2>&1
done
for w in 2 4; do
  TWG_RING_MODE=1 TWG_RING_WAVES=$w timeout 600 python bench.py --parts amips_ring --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_mode1_w$w.log 2>&1
done
timeout 300 python scripts/wbuild_time.py > gpurun_out/${TAG}_build_time.log 2>&1; cat gpurun_out/${TAG}_build_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_wbuild_launches.csv python scripts/wbuild_time.py --once > gpurun_out/${TAG}_wbuild_ncu.log 2>&1
for b in 16 20 24; do
  TWG_SORT_BITS=$b timeout 600 python bench.py --parts envelope --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_sortbits$b.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s13_ring_*.log')) + sorted(glob.glob('gpurun_out/r2s13_sortbits*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('/')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], d.get('extra', {}).get('decision_mismatches_vs_oracle_100k_sample'), d['roofline'].get('hbm_frac'))
PY
