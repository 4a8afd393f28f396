#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
NSIG_ADAM_TMA=2 timeout 600 $CS python tests/_variant_probe.py adam > gpurun_out/memcheck_adam.log 2>&1; grep -v "^=========     Host Frame\|^=========         Host" gpurun_out/memcheck_adam.log | head -40
