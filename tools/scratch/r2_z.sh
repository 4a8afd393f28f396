#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_render_rays --launch-skip 1 --launch-count 1 -o gpurun_out/r02_render -f python tools/bench_render.py --views 2 > gpurun_out/ncu_render.log 2>&1
tail -2 gpurun_out/ncu_render.log | cut -c1-300; ls -la gpurun_out/r02_render.ncu-rep
