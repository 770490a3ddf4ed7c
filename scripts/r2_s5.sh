#!/bin/bash
# round 2, session 5: gpu tests; nearest: round-scheduled lanes vs packets; envelope step: oriented bound on/off, own sort timing
TAG=r2s5
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
run_near() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --parts nearest --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_nearest_$name.log 2>&1
  python - <<PY
import json
for l in open('gpurun_out/${TAG}_nearest_$name.log'):
    if l.startswith('{'):
        d=json.loads(l); print('nearest $name', '%.3e pts/s'%d['value'], '%.2f ms'%d['ms_per_step'], d['extra']['parity_vs_brute_force'])
PY
}
run_near rounds64 TWG_NEAREST_MODE=1
run_near rounds32 TWG_NEAREST_MODE=1 TWG_NEAREST_GROUP=32
run_near rounds256 TWG_NEAREST_MODE=1 TWG_NEAREST_GROUP=256
run_near rounds_q8 TWG_NEAREST_MODE=1 TWG_ENV_QUORUM=8
run_near rounds_q24 TWG_NEAREST_MODE=1 TWG_ENV_QUORUM=24
run_near packet TWG_NEAREST_MODE=2 TWG_NEAREST_BUDGET=1000000
for B in 1 0; do
TWG_ENV_BOUND=$B timeout 600 python bench.py --parts envelope --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_env_bound$B.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/${TAG}_env_bound$B.log'):
    if l.startswith('{'):
        d=json.loads(l); print('envelope bound=$B', '%.3e pts/s'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['extra']['decision_mismatches_vs_oracle_100k_sample'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_env.csv python bench.py --parts envelope --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nearest_rounds -c 1 -o gpurun_out/${TAG}_near python bench.py --parts nearest --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_near.log 2>&1
ls gpurun_out/${TAG}* | head -30
