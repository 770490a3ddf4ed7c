#!/bin/bash
# round 2, session 19 (4 GPUs): multi-device context on 4 real devices, the driver's own torchrun launch at N=4 with the default parts
TAG=r2s19
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt; free -g | head -2 >> gpurun_out/${TAG}_gpus.txt
./tests/_build/test_multi 4 1 > gpurun_out/${TAG}_cpp_multi.log 2>&1; tail -5 gpurun_out/${TAG}_cpp_multi.log
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3) > gpurun_out/${TAG}_bench_n4.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2s19_bench_n4.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('envelope weak N=%d' % d['n_gpus'], '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
        for k,p in d['parts'].items(): print(k, p['scaling'], '%.3e'%p['value'], '%.3f ms'%p['ms_per_step'], 'e2e %.3e'%p['e2e']['value'])
PY
tail -3 gpurun_out/${TAG}_bench_n4.log | cut -c1-300
