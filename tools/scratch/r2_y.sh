#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optin_variants_gpu.py -x -q -m gpu 2>&1 | tail -12
