cd $GRAFT_REPO_ROOT
timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"
NSIG_NO_PDL=1 timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"
timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -3
