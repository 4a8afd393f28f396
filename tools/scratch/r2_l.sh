#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
for c in 1 2 3 8; do echo "== adam ctas/sm $c"; NSIG_ADAM_CTAS_PER_SM=$c step; done
done
NSIG_ADAM_CTAS_PER_SM=1 timeout 300 python tools/timeline_step.py --replays 1 --out gpurun_out/r02_timeline_step_adam1.txt > /dev/null 2>&1
NSIG_ADAM_CTAS_PER_SM=2 timeout 300 python tools/timeline_step.py --replays 1 --out gpurun_out/r02_timeline_step_adam2.txt > /dev/null 2>&1
timeout 600 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -3
