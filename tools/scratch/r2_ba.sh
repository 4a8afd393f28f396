#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
echo "== TC NO_SIDE"; NSIG_DEC_TC=1 NSIG_DEC_NO_SIDE=1 NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py 2>&1 | tail -12
echo "== TC default"; NSIG_DEC_TC=1 NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py 2>&1 | tail -12
