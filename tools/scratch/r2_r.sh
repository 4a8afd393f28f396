#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_dec_conv|k_dec_wgrad|k_dec_head" --launch-skip 41 --launch-count 32 -o gpurun_out/r02_decoder -f python tools/bench_decoder.py --iters 1 > gpurun_out/ncu_dec.log 2>&1
tail -3 gpurun_out/ncu_dec.log
ls -la gpurun_out/r02_decoder.ncu-rep
