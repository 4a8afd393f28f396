#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_raymarching_gpu.py tests/test_ref_cuda_parity.py tests/test_full_size_gpu.py tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -8
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
echo "== 3-pass march"; NSIG_MARCH_3PASS=1 step
echo "== fused march"; step
done
timeout 300 python tools/graph_offsets.py --out gpurun_out/r02_graph_offsets_fusedmarch.txt 2>&1 | tail -19
timeout 300 python tools/bench_render.py 2>&1 | tail -5
