#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -k "premarched or march_ahead or graph_step" 2>&1 | tail -15
for ma in 0 1 0 1; do
  echo "== march-ahead $ma"; timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 --march-ahead $ma 2> gpurun_out/ai_$ma.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'fwd', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'launches', d['gpu_launches'])
"; tail -3 gpurun_out/ai_$ma.err
done
