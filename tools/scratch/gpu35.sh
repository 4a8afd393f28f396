cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -2
timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' > gpurun_out/bench_v9.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v9.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
