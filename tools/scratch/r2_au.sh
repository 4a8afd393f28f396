#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
NSIG_DEC_TC=1 NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py > gpurun_out/dec_trace_tc.txt 2>&1; cat gpurun_out/dec_trace_tc.txt | tail -22
