#!/bin/bash
# call J: decoder tests (new determinism/prepared test), wgrad variants after the tail fix, same-box A/B of the step
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_train_step_gpu.py -x -q -m gpu 2>&1 | tail -3
for cfg in "1 1 16" "0 1 16" "0 2 16" "0 3 16" "0 4 16" "0 3 8" "0 3 12"; do
  set -- $cfg
  echo "== atomic=$1 side_streams=$2 G=$3"
  NSIG_DEC_WGRAD_ATOMIC=$1 NSIG_DEC_SIDE_STREAMS=$2 NSIG_DEC_WGRAD_G=$3 timeout 200 python tools/bench_decoder.py 2>&1 | tail -2 | head -1
done
step() { timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
  echo "== step rep $rep: atomic 1 stream, no prep (round-2 baseline)"; NSIG_NO_DEC_PREP=1 NSIG_DEC_WGRAD_ATOMIC=1 NSIG_DEC_SIDE_STREAMS=1 step
  echo "== step rep $rep: atomic 1 stream, prep";  NSIG_DEC_WGRAD_ATOMIC=1 NSIG_DEC_SIDE_STREAMS=1 step
  echo "== step rep $rep: two-phase 3 streams, prep"; NSIG_DEC_SIDE_STREAMS=3 step
  echo "== step rep $rep: two-phase 2 streams, prep"; NSIG_DEC_SIDE_STREAMS=2 step
  echo "== step rep $rep: two-phase 4 streams, prep"; NSIG_DEC_SIDE_STREAMS=4 step
done
