#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -15
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
echo "== no lookahead"; NSIG_NO_LOOKAHEAD=1 step
for c in 1 2 4; do echo "== lookahead, adam ctas/sm $c"; NSIG_ADAM_CTAS_PER_SM=$c step; done
done
echo "== lookahead, adam 2, sum-ahead 4"; NSIG_ADAM_CTAS_PER_SM=2 NSIG_SUM_AHEAD_CTAS_PER_SM=4 step
echo "== lookahead, adam 2, sum-ahead 1"; NSIG_ADAM_CTAS_PER_SM=2 NSIG_SUM_AHEAD_CTAS_PER_SM=1 step
NSIG_ADAM_CTAS_PER_SM=2 timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_lookahead.txt 2>&1 | tail -24
