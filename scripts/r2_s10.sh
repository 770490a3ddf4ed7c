#!/bin/bash
# round 2, session 10 (2 GPUs): multi-device context on real devices (Python + C++), torchrun bench N=2 incl. strong-scaling parts
TAG=r2s10
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt
nvidia-smi topo -m >> gpurun_out/${TAG}_gpus.txt 2>&1
(time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x) > gpurun_out/${TAG}_pytest_multi.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_multi.log
./tests/_build/test_multi 2 1 > gpurun_out/${TAG}_cpp_multi.log 2>&1; cat gpurun_out/${TAG}_cpp_multi.log
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --parts envelope,winding,envelope_strong,winding_strong) > gpurun_out/${TAG}_bench_n2.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s10_bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('envelope weak N=2', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
        for k,p in d['parts'].items(): print(k, p['scaling'], '%.3e'%p['value'], '%.3f ms'%p['ms_per_step'], 'e2e %.3e'%p['e2e']['value'], {a:b for a,b in p['extra'].items() if 'mism' in a or 'gbs' in a})
PY
(time timeout 900 python bench.py --parts envelope,winding --steps 5 --warmup 3 --no-cpu) > gpurun_out/${TAG}_bench_n1.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s10_bench_n1.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('envelope N=1', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
        for k,p in d['parts'].items(): print(k, '%.3e'%p['value'], 'e2e %.3e'%p['e2e']['value'])
PY
