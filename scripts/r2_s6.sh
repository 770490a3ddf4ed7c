#!/bin/bash
# round 2, session 6: nearest work counters; one-ring kernel cp.async A/B; amips tests
TAG=r2s6
mkdir -p gpurun_out
timeout 600 python scripts/near_diag.py 2000000 > gpurun_out/${TAG}_near_diag.log 2>&1
cat gpurun_out/${TAG}_near_diag.log
(timeout 900 python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_smoothing_pass.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_amips.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_amips.log
for A in 1 0; do for W in 3 2; do
TWG_RING_ASYNC=$A TWG_RING_WAVES=$W timeout 600 python bench.py --parts amips_ring --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_ring_a${A}_w${W}.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/${TAG}_ring_a${A}_w${W}.log'):
    if l.startswith('{'):
        d=json.loads(l); print('ring async=$A waves=$W', '%.3e tets/s'%d['value'], '%.3f ms'%d['ms_per_step'], 'hbm frac %.3f'%d['roofline']['hbm_frac'], 'e2e %.3e'%d['e2e']['value'], d['extra']['parity_vs_oracle'])
PY
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:amips_ring_async -c 1 -o gpurun_out/${TAG}_ring python bench.py --parts amips_ring --steps 1 --warmup 3 --no-cpu --scale 0.32 > gpurun_out/${TAG}_ncu_ring.log 2>&1
ls gpurun_out/${TAG}* | head
