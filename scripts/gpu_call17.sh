#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_envelope.py tests/test_gpu_smoothing_pass.py tests/test_gpu_winding.py -m gpu -x -q) > gpurun_out/s17_pytest.log 2>&1
tail -3 gpurun_out/s17_pytest.log
for q in 16 24 8; do echo "quorum=$q $(TWG_ENV_QUORUM=$q python scripts/env_ab.py 2>&1 | tail -1 | cut -c1-700)"; done > gpurun_out/s17_env.log
cat gpurun_out/s17_env.log
python bench.py --parts envelope_faces --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/s17_faces.log; python scripts/bench_summary.py gpurun_out/s17_faces.log | grep envelope
python scripts/prof_part.py winding 4e6 3 2>&1 | tail -1
