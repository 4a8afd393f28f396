#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
V=/root/repo/tools/scratch/libs/libnsig_pospf.so
for i in 1 2; do
echo "== default (rolled)"; timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-120
echo "== + position prefetch"; NSIG_LIB=$V timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-120
done
rend() { timeout 300 python tools/bench_render.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(v['fused_whole_frame']['ms_per_frame']) for k,v in d.items()})"; }
echo "== render default"; rend; echo "== render prefetch"; NSIG_LIB=$V rend
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; }
for rep in 1 2; do echo "== step default"; step; echo "== step prefetch"; NSIG_LIB=$V step; done
NSIG_LIB=$V timeout 600 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py -x -q -m gpu 2>&1 | tail -2
