#!/bin/bash
# winding diet check: parity tests + timing + ncu; new bench parts (faces, resident ring e2e); C++ adapters incl. TetMesh
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_winding.py tests/test_gpu_mesh.py tests/test_cpp_adapters.py -m gpu -x -q) > gpurun_out/s3_pytest.log 2>&1
tail -3 gpurun_out/s3_pytest.log
python scripts/prof_part.py winding 4e6 5 > gpurun_out/s3_time_wind.log 2>&1; tail -1 gpurun_out/s3_time_wind.log
python bench.py --parts envelope_faces,amips_ring --steps 3 --warmup 3 > gpurun_out/s3_bench_parts.log 2>&1; tail -c 3000 gpurun_out/s3_bench_parts.log
ncu --set full --clock-control none --import-source on -k regex:winding_kernel -s 1 -c 1 -f -o gpurun_out/s3_wind python scripts/prof_part.py winding 2e6 2 > gpurun_out/s3_ncu_wind.log 2>&1
