#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_decoder.py 2>&1 | tail -2; }
run A=1
run A=1
for v in "A=1" "A=1"; do
echo "== step $v"; env $v timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/bf.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'launches', d['gpu_launches'])
"; done
timeout 600 python tools/graph_offsets.py --out gpurun_out/off_tc.txt > /dev/null 2>gpurun_out/off1.err; tail -2 gpurun_out/off1.err; cat gpurun_out/off_tc.txt
