cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -x -q 2>&1 | grep -v Warning | tail -12
for mode in overlap merged split; do
timeout 300 python bench.py --no-cpu-baseline --no-render --render-mode $mode 2>&1 | grep '^{' > gpurun_out/bench_$mode.json; python -c "
import json; d=json.load(open('gpurun_out/bench_$mode.json')); print('$mode :', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['avg_launch_ms'], d['roofline']['samples_per_launch'])"
done
