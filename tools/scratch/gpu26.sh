cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-render > gpurun_out/bench_n$n.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/bench_n$n.log > gpurun_out/bench_n$n.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print('N=$n :', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['exchange'])" || tail -15 gpurun_out/bench_n$n.log
done
