#!/bin/bash
N=$1
cd /root/repo; mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_exchange_gpu.py -x -q -m gpu 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-render > gpurun_out/r02_bench_n${N}_v5.json 2> gpurun_out/r02_bench_n${N}_v5.err; tail -3 gpurun_out/r02_bench_n${N}_v5.err | cut -c1-200
python - <<P
import json
for l in open('gpurun_out/r02_bench_n${N}_v5.json'):
    try: d=json.loads(l)
    except Exception: continue
    print('N=${N} ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    c=d.get('configs4_shard262144',{}); print('configs4 ms', c.get('ms_per_step'), 'value', c.get('value'))
    for k in ('grad_check_1_vs_n','exchange_check'):
        print(k, json.dumps(d.get(k))[:500])
P
