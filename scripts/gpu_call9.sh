#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_envelope.py -m gpu -x -q) > gpurun_out/s9_pytest.log 2>&1
tail -3 gpurun_out/s9_pytest.log
python bench.py --parts envelope_faces --steps 3 --warmup 3 --no-cpu > gpurun_out/s9_bench_faces.log 2>&1; python scripts/bench_summary.py gpurun_out/s9_bench_faces.log
python scripts/prof_part.py faces 100000 3 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:env_faces -s 3 -c 1 -f -o gpurun_out/s9_faces python bench.py --parts envelope_faces --steps 1 --warmup 3 --no-cpu --scale 0.25 > gpurun_out/s9_ncu_faces.log 2>&1
