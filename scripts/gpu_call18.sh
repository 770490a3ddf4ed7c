#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_envelope.py tests/test_cpp_adapters.py -m gpu -x -q) > gpurun_out/s18_pytest.log 2>&1
tail -3 gpurun_out/s18_pytest.log
python bench.py --parts envelope_faces --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/s18_faces.log; python scripts/bench_summary.py gpurun_out/s18_faces.log | grep envelope; grep -o '"decision_mismatches[^,]*' gpurun_out/s18_faces.log
python scripts/prof_part.py faces 100000 3 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:env_faces -s 3 -c 1 -f -o gpurun_out/s18_faces python bench.py --parts envelope_faces --steps 1 --warmup 3 --no-cpu --scale 0.25 > gpurun_out/s18_ncu_faces.log 2>&1
