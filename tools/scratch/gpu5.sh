cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/scratch/ar_bench.py 2>&1 | grep -v "^W1017\|^\*\*\*" | tail -30
