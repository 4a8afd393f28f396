#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
V=/root/repo/tools/scratch/libs/libnsig_v2.so
for i in 1 2; do
echo "== default (v1 rolled)"; timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-120
echo "== v2 pair gather"; NSIG_LIB=$V timeout 300 python tools/bench_field.py --rays 8704 2>&1 | tail -1 | cut -c1-120
done
rend() { timeout 300 python tools/bench_render.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(v['fused_whole_frame']['ms_per_frame']) for k,v in d.items()})"; }
echo "== render default"; rend; echo "== render v2"; NSIG_LIB=$V rend
