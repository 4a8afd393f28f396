#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for i in 1 2; do
echo "== default"; timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-400
echo "== rolled fwd"; NSIG_LIB=/root/repo/tools/scratch/libs/libnsig_fwdrolled.so timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-400
done
NSIG_LIB=/root/repo/tools/scratch/libs/libnsig_fwdrolled.so timeout 600 python -m pytest tests/test_field_gpu.py tests/test_hash_gpu.py -x -q -m gpu 2>&1 | tail -2
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; }
for rep in 1 2; do echo "== step default"; step; echo "== step rolled"; NSIG_LIB=/root/repo/tools/scratch/libs/libnsig_fwdrolled.so step; done
