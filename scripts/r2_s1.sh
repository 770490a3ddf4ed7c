#!/bin/bash
# round 2, session 1: all gpu tests on the refactored library + per-kernel breakdown of the envelope / nearest steps at full size
TAG=r2s1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_env.csv python bench.py --parts envelope,nearest --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/${TAG}_bench_under_ncu.log
