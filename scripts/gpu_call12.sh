#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/s12_pytest_gpu.log 2>&1
tail -3 gpurun_out/s12_pytest_gpu.log
python bench.py --parts nearest,amips_quality --steps 3 --warmup 3 > gpurun_out/s12_bench_parts.log 2>&1; python scripts/bench_summary.py gpurun_out/s12_bench_parts.log
