#!/bin/bash
# GPU call: mesh tests + ncu captures of the current kernels (scratch driver; outputs under gpurun_out/)
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_mesh.py -m gpu -x -q) > gpurun_out/s2_pytest_mesh.log 2>&1
tail -3 gpurun_out/s2_pytest_mesh.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:env_points -s 1 -c 1 -f -o gpurun_out/s2_env  python scripts/prof_part.py envelope 10e6 2 > gpurun_out/s2_ncu_env.log 2>&1
$NCU -k regex:amips_ring -s 1 -c 1 -f -o gpurun_out/s2_ring python scripts/prof_part.py ring 16e6 2 > gpurun_out/s2_ncu_ring.log 2>&1
$NCU -k regex:winding_kernel -s 1 -c 1 -f -o gpurun_out/s2_wind python scripts/prof_part.py winding 2e6 2 > gpurun_out/s2_ncu_wind.log 2>&1
$NCU -k regex:nearest_kernel -s 0 -c 1 -f -o gpurun_out/s2_near python scripts/prof_part.py nearest 2e6 1 > gpurun_out/s2_ncu_near.log 2>&1
$NCU -k regex:"mesh_|amips_ring" -s 3 -c 3 -f -o gpurun_out/s2_mesh python scripts/prof_part.py mesh 120 2 > gpurun_out/s2_ncu_mesh.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.1 --no-cpu > gpurun_out/s2_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
