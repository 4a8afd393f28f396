set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in merged overlap split; do
  timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render --render-mode $mode > gpurun_out/r02_bench_h_$mode.json 2> /dev/null; python -c "
import json,sys
for l in open('gpurun_out/r02_bench_h_$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$mode', 'default', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
for c in 2 3; do
  NSIG_FIELD_CTAS_PER_SM=$c timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render --render-mode overlap > gpurun_out/r02_bench_h_overlap$c.json 2> /dev/null; python -c "
import json,sys
for l in open('gpurun_out/r02_bench_h_overlap$c.json'):
    if l.startswith('{'):
        d=json.loads(l); print('overlap', 'field ctas/sm $c', d['ms_per_step'], d['e2e']['ms_per_step'])"
  NSIG_FIELD_CTAS_PER_SM=$c timeout 300 python bench.py --no-extra --no-cpu-baseline --no-render --render-mode merged > gpurun_out/r02_bench_h_merged$c.json 2> /dev/null; python -c "
import json,sys
for l in open('gpurun_out/r02_bench_h_merged$c.json'):
    if l.startswith('{'):
        d=json.loads(l); print('merged', 'field ctas/sm $c', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
