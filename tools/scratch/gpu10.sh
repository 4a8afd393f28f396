cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_train_step_gpu.py -x -q 2>&1 | grep -v Warning | tail -40
