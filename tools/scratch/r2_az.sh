#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for v in "A=1" "NSIG_DEC_WGRAD_LATE=1" "NSIG_DEC_NO_SIDE=1"; do
echo "== TC $v"; env NSIG_DEC_TC=1 $v NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py 2>&1 | tail -9 | cut -c1-12,100-140
done
