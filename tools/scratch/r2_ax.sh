#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_decoder.py 2>&1 | tail -2 | head -1; }
run A=1
run NSIG_DEC_TC=1
run NSIG_DEC_TC=1 NSIG_DEC_SIDE_STREAMS=2
run NSIG_DEC_TC=1 NSIG_DEC_SIDE_STREAMS=4
run NSIG_DEC_TC=1 NSIG_DEC_WGRAD_LATE=1
run NSIG_DEC_WGRAD_LATE=1
run NSIG_DEC_SIDE_STREAMS=2
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -2
