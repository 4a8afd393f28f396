#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/dc.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'launches', d['gpu_launches'])
"; grep -v "Warning\|detach\|return float" gpurun_out/dc.err | tail -3
