#!/bin/bash
# round 2, session 2: gpu tests, bench smoke at scale 0.1, nearest packet kernel A/B at full size + ncu capture
TAG=r2s2
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
(time timeout 900 python bench.py --scale 0.1 --steps 3 --warmup 3 --cpu-budget 2) > gpurun_out/${TAG}_bench_scale0.1.log 2>&1
tail -c 400 gpurun_out/${TAG}_bench_scale0.1.log
(time timeout 600 python bench.py --parts nearest --steps 5 --warmup 3 --no-cpu) > gpurun_out/${TAG}_nearest_packet.log 2>&1
TWG_NEAREST_MODE=0 timeout 600 python bench.py --parts nearest --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_nearest_lane.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nearest_packet -c 1 -o gpurun_out/${TAG}_near python bench.py --parts nearest --steps 1 --warmup 3 --no-cpu --scale 0.2 > gpurun_out/${TAG}_ncu_near.log 2>&1
ls -la gpurun_out/${TAG}*
