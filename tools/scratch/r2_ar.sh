#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
NSIG_LIB=tools/scratch/libs/libnsig_trace.so timeout 300 python tools/dec_trace.py > gpurun_out/dec_trace.txt 2>&1; cat gpurun_out/dec_trace.txt | tail -30
