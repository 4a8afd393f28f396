#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 200 python tools/bench_decoder.py 2>&1 | tail -2; done
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do echo "== step"; step; done
