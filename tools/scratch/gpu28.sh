cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"
timeout 120 python tools/bench_decoder.py --B 48 2>&1 | grep "^decoder"
