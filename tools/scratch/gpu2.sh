cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu2.log
echo "== bench_field half2"; timeout 300 python tools/bench_field.py 2>&1 | tail -1
echo "== bench_field fp32"; NSIG_FP32_TABLES=1 timeout 300 python tools/bench_field.py 2>&1 | tail -1
