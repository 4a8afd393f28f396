#!/bin/bash
# call I: decoder backward variants (side streams x reduction x groups), parity tests, then the step
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu 2>&1 | tail -3
for atomic in 0 1; do for ns in 1 2 3; do for g in 16 32; do
  echo "== atomic=$atomic side_streams=$ns G=$g"
  NSIG_DEC_WGRAD_ATOMIC=$atomic NSIG_DEC_SIDE_STREAMS=$ns NSIG_DEC_WGRAD_G=$g timeout 200 python tools/bench_decoder.py 2>&1 | tail -2
done; done; done
for ns in 1 2 3; do
  echo "== step, side_streams=$ns"
  NSIG_DEC_SIDE_STREAMS=$ns timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
done
echo "== step, atomic, side_streams=2"
NSIG_DEC_WGRAD_ATOMIC=1 NSIG_DEC_SIDE_STREAMS=2 timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
