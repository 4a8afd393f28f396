#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py > gpurun_out/dec_trace2.txt 2>&1; cat gpurun_out/dec_trace2.txt | tail -22
timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/as.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
"
