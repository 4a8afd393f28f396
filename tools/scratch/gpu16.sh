cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -12
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_fl.json; python -c "
import json; d=json.load(open('gpurun_out/bench_fl.json')); print('default :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['avg_launch_ms'])"
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 300 python tools/profile_step.py --out gpurun_out/profile_step.txt > gpurun_out/profile_step.log 2>&1; head -12 gpurun_out/profile_step.txt
