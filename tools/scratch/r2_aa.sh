#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/bench_render.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(v['fused_whole_frame']['ms_per_frame'], v['fused_staged_4096']['ms_per_frame']) for k,v in d.items()})"; done
