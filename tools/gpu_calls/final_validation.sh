#!/bin/bash
# final validation + evidence with the final tree
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v6.json 2> gpurun_out/r02_bench_n1_v6.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_v6.json').read().strip().splitlines()[-1])
print('N=1 ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], [(o['kernel'][:20], round(o['frac'],3)) for o in d['roofline_other']])
print('render', {k:v['ms_per_frame'] for k,v in d['render'].items()})
print('c2', d['configs2_360_wtmk']['ms_per_step'], 'c4', d['configs4_shard262144']['ms_per_step'], 'refcuda', d['ref_cuda']['ms_per_step'], d['ref_cuda']['ours_e2e_over_ref_cuda'])
print('cpu', d['cpu_baseline']['value'], d['clocks'])
P
timeout 300 python tools/profile_step.py --out gpurun_out/r02_step_breakdown_v3.txt > gpurun_out/profile_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-extra --no-graph > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-120
