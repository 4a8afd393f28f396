#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_decoder.py 2>&1 | tail -2 | head -1; }
run A=1
run NSIG_DEC_TC=1
run NSIG_DEC_TC=1 NSIG_DEC_SIDE_STREAMS=2
run NSIG_DEC_TC=1 NSIG_DEC_SIDE_STREAMS=4
run NSIG_DEC_TC=1 NSIG_DEC_WGRAD_LATE=1
run A=1
run NSIG_DEC_TC=1
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
for tc in 0 1 0 1; do
echo "== step NSIG_DEC_TC=$tc"; NSIG_DEC_TC=$tc timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/ay.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
"; done
