#!/bin/bash
# The round's evidence pass on one B200 (about 20 minutes): every GPU test, smoke, sanitizers, the N=1 bench as the driver runs it
# (both arms), the launch list under ncu, one ncu --set full capture per dominant kernel. TAG names the outputs (gpurun_out/TAG_*).
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt; nproc >> gpurun_out/${TAG}_gpu.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
for tool in memcheck synccheck racecheck; do
  (time timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python __graft_entry__.py --smoke) > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_$tool.log | tail -1
done
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench_n1.log 2>&1
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/${TAG}_bench_reference_arm.log 2>&1
# launch list: every part but the two that are thousands of tiny launches (pass_stream) or a 100 M one-shot (winding_oneshot)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --scale 0.1 --no-cpu \
  --parts envelope,envelope_faces,envelope_faces_c1,nearest,amips,amips_literal,amips_quality,amips_ring,winding > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
bash scripts/r2_profile.sh ${TAG}
python scripts/latency.py > gpurun_out/${TAG}_latency.log 2>&1
python scripts/near_diag.py 2000000 > gpurun_out/${TAG}_near_diag.log 2>&1
tail -c 300 gpurun_out/${TAG}_bench_n1.log
