cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-render > gpurun_out/bench_n2_final.log 2>&1; echo rc=$?
grep '^{' gpurun_out/bench_n2_final.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['exchange'])" || tail -5 gpurun_out/bench_n2_final.log
