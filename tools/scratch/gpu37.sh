cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; echo "rc=$? wall $(( $(date +%s) - S )) s"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_final.log') if l.startswith('{')][-1]); print(d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['sample'][-60:], d['roofline']['frac'], d['gpu_launches'])"
