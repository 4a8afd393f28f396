cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_train_step_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], 'e2e', d['e2e'], d['gpu_launches'])"
