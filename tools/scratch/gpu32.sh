cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_render_gpu.py tests/test_field_gpu.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/bench_v7.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v7.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_frame'] for k,v in d['render'].items()})"
export NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_field_bwd' -f -o gpurun_out/bwd_v6 python tools/ncu_step.py > gpurun_out/ncu_bwd.log 2>&1
python tools/ncu_summary.py gpurun_out/bwd_v6.ncu-rep > gpurun_out/bwd_v6.txt
ncu -i gpurun_out/bwd_v6.ncu-rep --page source --csv > gpurun_out/bwd_v6_source.csv 2>/dev/null
rm -f gpurun_out/bwd_v6.ncu-rep
