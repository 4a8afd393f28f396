cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py tests/test_grid_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -3
timeout 120 python tools/bench_field.py 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('field fwd/bwd ms', d['field_fwd_ms'], d['field_bwd_ms'])"
timeout 300 python bench.py --no-cpu-baseline --no-render 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['field_backward_avg_launch_ms'])"
