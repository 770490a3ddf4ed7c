#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_envelope.py tests/test_cpp_adapters.py -m gpu -x -q) > gpurun_out/s10_pytest.log 2>&1
tail -3 gpurun_out/s10_pytest.log
python bench.py --parts envelope_faces --steps 3 --warmup 3 --no-cpu > gpurun_out/s10_bench_faces.log 2>&1; python scripts/bench_summary.py gpurun_out/s10_bench_faces.log | grep envelope
python scripts/prof_part.py faces 100000 3 2>&1 | tail -1
for mb in 6 8; do TWG_ENV_MINB=$mb python scripts/env_ab.py 2>&1 | tail -1; done > gpurun_out/s10_env_ab.log; cat gpurun_out/s10_env_ab.log
