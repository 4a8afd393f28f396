#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
export NSIG_LIB=/root/repo/tools/scratch/libs/libnsig_dec512.so
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 200 python tools/bench_decoder.py 2>&1 | tail -2; done
unset NSIG_LIB
for i in 1; do timeout 200 python tools/bench_decoder.py 2>&1 | tail -2; done
