#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_envelope.py tests/test_cpp_adapters.py -m gpu -x -q) > gpurun_out/s7_pytest.log 2>&1
tail -3 gpurun_out/s7_pytest.log
python bench.py --parts envelope_faces --steps 3 --warmup 3 > gpurun_out/s7_bench_faces.log 2>&1; python scripts/bench_summary.py gpurun_out/s7_bench_faces.log
python scripts/prof_part.py peaks 1 1 2>&1 | grep fp64 > gpurun_out/s7_peaks.log; cat gpurun_out/s7_peaks.log
ncu --set full --clock-control none --import-source on -k regex:env_faces -s 1 -c 1 -f -o gpurun_out/s7_faces python scripts/prof_part.py faces 100000 2 > gpurun_out/s7_ncu_faces.log 2>&1
tail -1 gpurun_out/s7_ncu_faces.log
