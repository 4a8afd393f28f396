#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== old wgrad"; NSIG_LIB=tools/scratch/libs/libnsig_oldwgrad.so timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== new wgrad"; timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== new wgrad G=32"; NSIG_DEC_WGRAD_G=32 timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== new wgrad, 2 side streams"; NSIG_DEC_SIDE_STREAMS=2 timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== new wgrad, 1 side stream"; NSIG_DEC_SIDE_STREAMS=1 timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== no wgrad (diagnosis)"; NSIG_DEC_SKIP_WGRAD=1 timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== old"; NSIG_LIB=tools/scratch/libs/libnsig_oldwgrad.so timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
echo "== new"; timeout 300 python tools/bench_decoder.py 2>&1 | tail -3
for lib in tools/scratch/libs/libnsig_oldwgrad.so nerf_signature_b200/libnsig_b200.so; do
echo "== step with $lib"; NSIG_LIB=$lib timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/an.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
"; done
