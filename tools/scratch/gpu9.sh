cd $GRAFT_REPO_ROOT
for e in 0 1; do
NSIG_DECODER_NCHW=$e timeout 600 python bench.py --no-cpu-baseline --no-render 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('NCHW=$e ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
NSIG_DECODER_NCHW=0 timeout 300 python tools/profile_step.py --out gpurun_out/profile_step3.txt > /dev/null 2>&1; head -3 gpurun_out/profile_step3.txt
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3
