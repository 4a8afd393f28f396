#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_field_gpu.py tests/test_hash_gpu.py tests/test_render_gpu.py tests/test_e2e_ref_parity_gpu.py tests/test_grid_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_render_rays --launch-skip 1 --launch-count 1 -o gpurun_out/r02_render_v2 -f python tools/bench_render.py --views 2 > gpurun_out/ncu_render.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_field_fwd' -o gpurun_out/r02_fwd_rolled -f python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
