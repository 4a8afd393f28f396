#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
NSIG_DEC_TC=1 timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== TC default trace"; NSIG_DEC_TC=1 NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py 2>&1 | tail -21 | cut -c1-12,100-140
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_decoder.py 2>&1 | tail -2 | head -1; }
run A=1
run NSIG_DEC_TC=1
run NSIG_DEC_TC=1 NSIG_DEC_SIDE_STREAMS=2
run NSIG_DEC_TC=1 NSIG_DEC_WGRAD_LATE=1
run A=1
run NSIG_DEC_TC=1
