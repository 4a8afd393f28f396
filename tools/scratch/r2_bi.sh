#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_decoder.py 2>&1 | tail -2; }
run NSIG_DEC_PDL=0
run A=1
echo "== trace PDL"; NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py 2>&1 | head -21 | cut -c1-150
for v in "NSIG_DEC_PDL=0" "A=1" "NSIG_DEC_PDL=0" "A=1"; do
echo "== step $v"; env $v timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/bi.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'launches', d['gpu_launches'])
"; grep -v "Warning\|detach\|return float" gpurun_out/bi.err | tail -3; done
