#!/bin/bash
# round 2, session 14: ncu capture of the member-stream one-ring kernel
TAG=r2s14
mkdir -p gpurun_out
TWG_RING_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:amips_ring_flat -c 1 -f -o gpurun_out/${TAG}_ring_flat python scripts/prof_part.py ring 16000000 2 > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py rep gpurun_out/${TAG}_ring_flat.ncu-rep gpurun_out/${TAG}_ring_flat.txt
head -45 gpurun_out/${TAG}_ring_flat.txt
