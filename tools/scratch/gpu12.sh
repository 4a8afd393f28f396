cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_train_step_gpu.py tests/test_field_gpu.py -x -q 2>&1 | grep -v Warning | tail -15
for g in 32 16; do for ns in 0 1; do NSIG_DEC_WGRAD_G=$g NSIG_DEC_NO_SIDE=$ns timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"; done; done
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_side.json; python -c "
import json; d=json.load(open('gpurun_out/bench_side.json')); print('side streams ON :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_noside.json; python -c "
import json; d=json.load(open('gpurun_out/bench_noside.json')); print('side streams OFF:', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
timeout 300 python tools/profile_step.py --out gpurun_out/profile_step.txt > gpurun_out/profile_step.log 2>&1; head -30 gpurun_out/profile_step.txt
