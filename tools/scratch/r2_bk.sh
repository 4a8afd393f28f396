#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in "NSIG_DEC_JOIN_LATE=1" "A=1" "NSIG_DEC_JOIN_LATE=1" "A=1"; do
echo "== step $v"; env $v timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/bk.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'launches', d['gpu_launches'])
"; grep -v "Warning\|detach\|return float" gpurun_out/bk.err | tail -3; done
timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_v4.txt > /dev/null 2>&1; tail -8 gpurun_out/r02_graph_offsets_v4.txt
