#!/bin/bash
TAG=${1:-sx}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
(time python bench.py) > gpurun_out/${TAG}_bench_n1.log 2>&1
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.1 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_points -s 1 -c 1 -f -o gpurun_out/${TAG}_env python scripts/prof_part.py envelope 10e6 2 > gpurun_out/${TAG}_ncu_env.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:winding_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_wind python scripts/prof_part.py winding 2e6 2 > gpurun_out/${TAG}_ncu_wind.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_kernel -s 0 -c 1 -f -o gpurun_out/${TAG}_near python scripts/prof_part.py nearest 2e6 1 > gpurun_out/${TAG}_ncu_near.log 2>&1
ncu --set full --clock-control none -k regex:"mesh_quality" -s 1 -c 1 -f -o gpurun_out/${TAG}_quality python bench.py --parts amips_quality --steps 1 --warmup 3 --no-cpu --scale 0.3 > gpurun_out/${TAG}_ncu_quality.log 2>&1
