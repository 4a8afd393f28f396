#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_raymarching_gpu.py -x -q -m gpu -k "premarched or march_ahead or march" 2>&1 | tail -5
for spec in "0 2" "1 1" "1 2" "1 3" "0 2" "1 2"; do
  set -- $spec
  echo "== march-ahead $1 ctas/sm $2"; NSIG_MARCH_AHEAD_CTAS_PER_SM=$2 timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 --march-ahead $1 2> gpurun_out/ak.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'fwd', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'launches', d['gpu_launches'])
"; grep -v Warning gpurun_out/ak.err | tail -2
done
NSIG_MARCH_AHEAD_CTAS_PER_SM=2 timeout 600 python tools/graph_offsets.py --march-ahead 1 --out gpurun_out/off_ma1_lim2.txt > /dev/null 2>gpurun_out/off1.err; tail -2 gpurun_out/off1.err
