cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_exchange_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | grep -v Warning | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-render > gpurun_out/bench_n2.log 2>&1; grep '^{' gpurun_out/bench_n2.log > gpurun_out/bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2 :', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['exchange'])" || tail -20 gpurun_out/bench_n2.log
timeout 200 python bench.py --no-render --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/bench_n1.json;  python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('N=1 :', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | grep '^{' | cut -c1-200
