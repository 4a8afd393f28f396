set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_exchange_gpu.py tests/test_sharded_gpu.py -m gpu -q > gpurun_out/pytest_exchange_n$N.log 2>&1; echo "exchange rc=$?"; tail -5 gpurun_out/pytest_exchange_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r02_bench_n$N.json; tail -8 gpurun_out/r02_bench_n$N.err
