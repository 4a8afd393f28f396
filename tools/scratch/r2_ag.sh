#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
echo "== memcheck: TMA adam probe"; NSIG_ADAM_TMA=2 timeout 600 $CS python tests/_variant_probe.py adam 2>&1 | tail -4
echo "== memcheck: fused march probe"; NSIG_MARCH_FUSED=1 timeout 900 $CS python tests/_variant_probe.py march 2>&1 | tail -4
echo "== memcheck: decoder tests"; timeout 1200 $CS python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== memcheck: field tests (backward kernels + forward)"; timeout 1500 $CS python -m pytest tests/test_field_gpu.py -x -q -m gpu -k "agree or forward or message" 2>&1 | tail -4
echo "== memcheck: lookahead test"; timeout 900 $CS python -m pytest tests/test_train_step_gpu.py -x -q -m gpu -k lookahead 2>&1 | tail -4
echo "== memcheck: render tests"; timeout 1500 $CS python -m pytest tests/test_render_gpu.py -x -q -m gpu 2>&1 | tail -4
