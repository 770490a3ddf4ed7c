#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_amips.py -m gpu -x -q) > gpurun_out/s15_pytest.log 2>&1
tail -3 gpurun_out/s15_pytest.log
for t in 0 1; do TWG_AMIPS_TMA=$t python scripts/prof_part.py amips 50e6 6 2>&1 | tail -1; done
ncu --set full --clock-control none --import-source on -k regex:amips_soa_tma -s 1 -c 1 -f -o gpurun_out/s15_amips python scripts/prof_part.py amips 16e6 2 > gpurun_out/s15_ncu.log 2>&1
