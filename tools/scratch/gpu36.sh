cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/bench_v10.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v10.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_per_frame'],2) for k,v in d['render'].items()})"
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 200 python tools/profile_step.py --out gpurun_out/profile_step.txt > gpurun_out/profile_step.log 2>&1; grep "march" gpurun_out/profile_step.txt
