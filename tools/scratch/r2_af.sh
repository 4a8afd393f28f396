#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
echo "== register Adam"; step
for c in 2 3; do echo "== TMA Adam $c CTA/SM, G kept in L2"; NSIG_ADAM_TMA=$c step; done
done
NSIG_ADAM_TMA=2 timeout 200 python tools/scratch/adam_ab.py 2>&1 | tail -2
