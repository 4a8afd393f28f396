#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== tests with NSIG_DEC_TC=1"; NSIG_DEC_TC=1 timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== bench mma.sync"; timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
echo "== bench tcgen05"; NSIG_DEC_TC=1 timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
