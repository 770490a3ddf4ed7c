#!/bin/bash
# full GPU evidence pass: every gpu test, smoke, the N=1 bench (both arms), launch list under ncu, ncu captures of the dominant kernels; TAG names the outputs
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/${TAG}_bench_n1.log 2>&1
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_reference_arm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.1 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
bash scripts/r2_profile.sh ${TAG}
python scripts/latency.py > gpurun_out/${TAG}_latency.log 2>&1
python scripts/near_diag.py 2000000 > gpurun_out/${TAG}_near_diag.log 2>&1
tail -c 400 gpurun_out/${TAG}_bench_n1.log
