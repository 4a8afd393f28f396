#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -25
timeout 300 python tools/timeline_step.py --out gpurun_out/r02_timeline_step.txt > /dev/null 2> gpurun_out/timeline.err; tail -3 gpurun_out/timeline.err
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_e2e_ref_parity_gpu.py -x -q -m gpu 2>&1 | tail -5
