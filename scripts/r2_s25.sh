#!/bin/bash
# round 2, session 25: fixed tests; point-kernel options on the kd hierarchy (group 128 with / without the oriented bound); winding leaf size
TAG=r2s25
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_tiny_calls.py -m gpu -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --parts envelope,envelope_faces --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_env_$name.log 2>&1; }
run g128 TWG_ENV_GROUP=128
run g128b0 TWG_ENV_GROUP=128 TWG_ENV_BOUND=0
run g256 TWG_ENV_GROUP=256
run g128q24 TWG_ENV_GROUP=128 TWG_ENV_QUORUM=24
for lf in 32 128; do
  TWG_WINDING_LEAF=$lf timeout 600 python bench.py --parts winding --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_wleaf$lf.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s25_env_*.log')) + sorted(glob.glob('gpurun_out/r2s25_wleaf*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('r2s25_')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], d['extra'].get('decision_mismatches_vs_oracle_100k_sample'), [(k, '%.3f' % p['ms_per_step']) for k, p in d.get('parts', {}).items()], d['roofline'].get('pairs_per_query'))
PY
