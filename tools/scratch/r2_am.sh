#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 600 python bench.py --no-extra --no-render --no-cpu-baseline --steps 30 --warmup 5 2> gpurun_out/am.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'fwd', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'launches', d['gpu_launches'])
"; grep -v "Warning\|tensor.detach\|return float" gpurun_out/am.err | tail -2; }
run A=1
run NSIG_LOOKAHEAD=1 NSIG_LOOKAHEAD_AT=field NSIG_ADAM_TMA=1 NSIG_FIELD_FWD_CTAS_PER_SM=3
run NSIG_LOOKAHEAD=1 NSIG_LOOKAHEAD_AT=field NSIG_ADAM_TMA=2 NSIG_FIELD_FWD_CTAS_PER_SM=3
run NSIG_LOOKAHEAD=1 NSIG_LOOKAHEAD_AT=field NSIG_ADAM_TMA=1 NSIG_FIELD_FWD_CTAS_PER_SM=4
run NSIG_LOOKAHEAD=1 NSIG_LOOKAHEAD_AT=field NSIG_ADAM_CTAS_PER_SM=1 NSIG_FIELD_FWD_CTAS_PER_SM=3
run NSIG_FIELD_FWD_CTAS_PER_SM=3
NSIG_LOOKAHEAD=1 NSIG_LOOKAHEAD_AT=field NSIG_ADAM_TMA=1 NSIG_FIELD_FWD_CTAS_PER_SM=3 timeout 600 python tools/graph_offsets.py --out gpurun_out/off_shadow.txt > /dev/null 2>gpurun_out/off1.err; tail -2 gpurun_out/off1.err
