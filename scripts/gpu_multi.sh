#!/bin/bash
# N-GPU check on one box (gpurun --gpus N -- 'bash scripts/gpu_multi.sh N TAG'): twg_create_multi on N real devices, then the driver's own torchrun launch of
# bench.py with the default parts (weak headline + strong-scaling parts)
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus_n$N.txt; free -g | head -2 >> gpurun_out/${TAG}_gpus_n$N.txt
./tests/_build/test_multi $N 1 > gpurun_out/${TAG}_cpp_multi_n$N.log 2>&1; tail -4 gpurun_out/${TAG}_cpp_multi_n$N.log
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3) > gpurun_out/${TAG}_bench_n$N.log 2>&1
python - $N $TAG <<'PY'
import json, sys
for l in open('gpurun_out/%s_bench_n%s.log' % (sys.argv[2], sys.argv[1])):
    if l.startswith('{'):
        d = json.loads(l)
        print('envelope weak N=%d' % d['n_gpus'], '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'])
        for k, p in d['parts'].items(): print(k, p['scaling'], '%.3e' % p['value'], '%.3f ms' % p['ms_per_step'], 'e2e %.3e' % p['e2e']['value'])
PY
tail -3 gpurun_out/${TAG}_bench_n$N.log | cut -c1-200
