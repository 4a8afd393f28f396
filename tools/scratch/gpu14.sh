cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_raymarching_gpu.py tests/test_decoder_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | grep -v Warning | tail -15
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_fl.json; python -c "
import json; d=json.load(open('gpurun_out/bench_fl.json')); print('fused losses :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
timeout 300 python bench.py --no-cpu-baseline --no-render --torch-losses 2>&1 | grep '^{' > gpurun_out/bench_tl.json; python -c "
import json; d=json.load(open('gpurun_out/bench_tl.json')); print('torch losses :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
NSIG_TORCH_EPILOGUE=1 timeout 300 python bench.py --no-cpu-baseline --no-render --torch-losses 2>&1 | grep '^{' > gpurun_out/bench_tle.json; python -c "
import json; d=json.load(open('gpurun_out/bench_tle.json')); print('torch losses+epilogue :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 300 python tools/profile_step.py --out gpurun_out/profile_step.txt > gpurun_out/profile_step.log 2>&1; head -60 gpurun_out/profile_step.txt
