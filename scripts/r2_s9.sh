#!/bin/bash
# round 2, session 9: small-call latency with the zero-copy slab (on / off), call stream, full gpu tests
TAG=r2s9
mkdir -p gpurun_out
python scripts/latency.py > gpurun_out/${TAG}_latency_fast.log 2>&1; tail -c 900 gpurun_out/${TAG}_latency_fast.log; echo
TWG_FAST_CALLS=0 python scripts/latency.py > gpurun_out/${TAG}_latency_slow.log 2>&1; tail -c 900 gpurun_out/${TAG}_latency_slow.log; echo
timeout 600 python bench.py --parts pass_stream --steps 3 --warmup 1 --no-cpu > gpurun_out/${TAG}_stream.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s9_stream.log'):
    if l.startswith('{'):
        d=json.loads(l); e=d['extra']; print({k:e[k] for k in ('calls','mismatches','rebatched_by_kind_calls_per_s','call_by_call_calls_per_s','call_by_call_us_per_call')})
PY
(time timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
