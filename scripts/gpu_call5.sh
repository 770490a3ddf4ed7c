#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_amips.py tests/test_gpu_mesh.py tests/test_gpu_winding.py -m gpu -x -q) > gpurun_out/s5_pytest.log 2>&1
tail -3 gpurun_out/s5_pytest.log
for w in 3 6; do TWG_RING_WAVES=$w python scripts/prof_part.py ring 50e6 5 2>&1 | tail -1; done > gpurun_out/s5_ring.log
python scripts/prof_part.py mesh 120 5 2>&1 | tail -1 >> gpurun_out/s5_ring.log
cat gpurun_out/s5_ring.log
for m in 3 4; do TWG_WINDING_MINB=$m python scripts/prof_part.py winding 4e6 4 2>&1 | tail -1; done > gpurun_out/s5_wind.log
cat gpurun_out/s5_wind.log
ncu --set full --clock-control none --import-source on -k regex:amips_ring -s 1 -c 1 -f -o gpurun_out/s5_ring python scripts/prof_part.py ring 16e6 2 > gpurun_out/s5_ncu_ring.log 2>&1
