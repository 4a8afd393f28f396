#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
echo "== memcheck: decoder tests"; timeout 600 $CS python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | grep -v "Host Frame\|^=========  *Host" | tail -12
echo "== bench sanity"; timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/bq.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['config']['decoder'][:60])
"; grep -v "Warning\|detach\|return float" gpurun_out/bq.err | tail -3
