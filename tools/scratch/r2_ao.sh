#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_e2e_ref_parity_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== erff"; NSIG_LIB=tools/scratch/libs/libnsig_erff.so timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
echo "== A&S"; timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
echo "== erff"; NSIG_LIB=tools/scratch/libs/libnsig_erff.so timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
echo "== A&S"; timeout 300 python tools/bench_decoder.py 2>&1 | tail -2
for lib in tools/scratch/libs/libnsig_erff.so nerf_signature_b200/libnsig_b200.so tools/scratch/libs/libnsig_erff.so nerf_signature_b200/libnsig_b200.so; do
echo "== step with $lib"; NSIG_LIB=$lib timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/ao.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'))
"; done
