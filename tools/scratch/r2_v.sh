#!/bin/bash
# multi-GPU: N from the first argument
N=$1
cd /root/repo; mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_exchange_gpu.py -x -q -m gpu 2>&1 | tail -4; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_v3.json 2> gpurun_out/r02_bench_n${N}_v3.err
tail -c 700 gpurun_out/r02_bench_n${N}_v3.json; tail -3 gpurun_out/r02_bench_n${N}_v3.err
