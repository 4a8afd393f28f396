cd $GRAFT_REPO_ROOT
timeout 600 python tools/bench_ref_cuda.py --steps 5 --warmup 2 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
