#!/bin/bash
# round 2, session 24: kd facet order as the default: envelope / robustness / tiny tests, option sweep of the point kernel on the new hierarchy
TAG=r2s24
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_envelope.py tests/test_gpu_robustness.py tests/test_gpu_tiny_calls.py tests/test_gpu_multi.py -m gpu -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --parts envelope --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_env_$name.log 2>&1
}
run base TWG_TRACE=0
run group32 TWG_ENV_GROUP=32
run group128 TWG_ENV_GROUP=128
run front16 TWG_ENV_FRONT=16
run front64 TWG_ENV_FRONT=64
run quorum8 TWG_ENV_QUORUM=8
run quorum24 TWG_ENV_QUORUM=24
run top32 TWG_ENV_TOP=32
run top128 TWG_ENV_TOP=128
run bound0 TWG_ENV_BOUND=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s24_env_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print(f.split('_env_')[-1], '%.3f ms' % d['ms_per_step'], '%.3e' % d['value'], d['extra'].get('decision_mismatches_vs_oracle_100k_sample'))
PY
