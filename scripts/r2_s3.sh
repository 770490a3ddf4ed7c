#!/bin/bash
# round 2, session 3: gpu tests incl. multi-context, nearest budget sweep at full size, ncu of the hybrid kernel
TAG=r2s3
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
for B in 24 48 96 192 400; do
  TWG_NEAREST_BUDGET=$B timeout 600 python bench.py --parts nearest --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_nearest_b$B.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/${TAG}_nearest_b$B.log'):
    if l.startswith('{'):
        d=json.loads(l); print('budget $B', '%.3e pts/s'%d['value'], '%.2f ms'%d['ms_per_step'], d['extra']['parity_vs_brute_force'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nearest_packet -c 1 -o gpurun_out/${TAG}_near python bench.py --parts nearest --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_near.log 2>&1
ls gpurun_out/${TAG}*
