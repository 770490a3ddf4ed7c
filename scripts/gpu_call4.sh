#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_envelope.py -m gpu -x -q) > gpurun_out/s4_pytest.log 2>&1
tail -3 gpurun_out/s4_pytest.log
export ENV_AB_SKIP_NEAREST=1
for pol in 0 1; do for grp in 64 128; do TWG_ENV_POLICY=$pol TWG_ENV_GROUP=$grp python scripts/env_ab.py 2>&1 | tail -1; done; done > gpurun_out/s4_env_ab.log
cat gpurun_out/s4_env_ab.log
for w in 3 16; do TWG_RING_WAVES=$w python scripts/prof_part.py ring 50e6 5 2>&1 | tail -1; done > gpurun_out/s4_ring.log
python scripts/prof_part.py mesh 120 5 2>&1 | tail -1 >> gpurun_out/s4_ring.log
cat gpurun_out/s4_ring.log
ncu --set full --clock-control none -k regex:dfma_kernel -s 1 -c 1 -f -o gpurun_out/s4_dfma python scripts/prof_part.py peaks 1 1 > gpurun_out/s4_ncu_peaks.log 2>&1
tail -2 gpurun_out/s4_ncu_peaks.log
ncu --set full --clock-control none --import-source on -k regex:amips_ring -s 1 -c 1 -f -o gpurun_out/s4_ring python scripts/prof_part.py ring 16e6 2 > gpurun_out/s4_ncu_ring.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_points -s 1 -c 1 -f -o gpurun_out/s4_env python scripts/prof_part.py envelope 10e6 2 > gpurun_out/s4_ncu_env.log 2>&1
