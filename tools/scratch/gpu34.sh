cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_v8.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v8.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 300 python tools/profile_step.py --out gpurun_out/profile_step.txt > gpurun_out/profile_step.log 2>&1; grep "composite\|march" gpurun_out/profile_step.txt
