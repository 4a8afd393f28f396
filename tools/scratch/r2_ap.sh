#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_sharded_gpu.py tests/test_exchange_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-render > gpurun_out/ap_n2.json 2> gpurun_out/ap_n2.err; tail -3 gpurun_out/ap_n2.err
python - <<'P'
import json
for l in open('gpurun_out/ap_n2.json'):
    try: d=json.loads(l)
    except Exception: continue
    print('N=2 ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    for k in ('configs4_shard262144','grad_check_1_vs_n','exchange_check'):
        print(k, json.dumps(d.get(k))[:600])
P
