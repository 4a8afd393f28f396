#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
NSIG_ADAM_TMA=0 timeout 200 python tools/scratch/adam_ab.py 2>&1 | tail -2
for c in 1 2 3; do NSIG_ADAM_TMA=$c timeout 200 python tools/scratch/adam_ab.py 2>&1 | tail -2; done
timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -3
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
echo "== register Adam"; NSIG_ADAM_TMA=0 step
for c in 1 2 3; do echo "== TMA Adam $c CTA/SM"; NSIG_ADAM_TMA=$c step; done
done
NSIG_ADAM_TMA=2 timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_tma2.txt 2>&1 | sed -n 3,10p
