#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -5
for fs in 1 0 0 1; do
  echo "== NSIG_NO_FUSED_SUM=$fs"; NSIG_NO_FUSED_SUM=$fs timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/al.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'fwd', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'launches', d['gpu_launches'])
"; grep -v Warning gpurun_out/al.err | tail -2
done
timeout 600 python tools/graph_offsets.py --out gpurun_out/off_fusedsum.txt > /dev/null 2>gpurun_out/off1.err; tail -2 gpurun_out/off1.err
