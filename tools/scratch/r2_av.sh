#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== tests with NSIG_DEC_TC=1"; NSIG_DEC_TC=1 timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -6
echo "== tests default"; timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
NSIG_DEC_TC=1 NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py > gpurun_out/dec_trace_tc2.txt 2>&1; cat gpurun_out/dec_trace_tc2.txt | tail -20 | cut -c1-150
echo "== bench mma.sync"; timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
echo "== bench tcgen05"; NSIG_DEC_TC=1 timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
