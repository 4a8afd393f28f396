set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
for v in 0 1; do timeout 60 tools/scratch/tc_probe $v > gpurun_out/tc_probe_$v.log 2>&1; echo "probe $v rc=$?"; cat gpurun_out/tc_probe_$v.log; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log | cut -c1-300
NSIG_BWD_TC=1 timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log | cut -c1-300
NSIG_BWD_TC=0 timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_mma.log 2>&1; tail -n 1 gpurun_out/bench_field_mma.log
NSIG_BWD_TC=1 timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_tc.log 2>&1; tail -n 1 gpurun_out/bench_field_tc.log
timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_b_defer.json 2> gpurun_out/r02_bench_b_defer.err; tail -c 600 gpurun_out/r02_bench_b_defer.json; tail -3 gpurun_out/r02_bench_b_defer.err
timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render --no-defer > gpurun_out/r02_bench_b_nodefer.json 2> gpurun_out/r02_bench_b_nodefer.err; tail -c 600 gpurun_out/r02_bench_b_nodefer.json
NSIG_TORCH_SCALER=1 timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render --no-defer > gpurun_out/r02_bench_b_torchscaler.json 2> gpurun_out/r02_bench_b_torchscaler.err; tail -c 600 gpurun_out/r02_bench_b_torchscaler.json
NSIG_BWD_TC=1 timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_b_tc.json 2> gpurun_out/r02_bench_b_tc.err; tail -c 600 gpurun_out/r02_bench_b_tc.json; tail -3 gpurun_out/r02_bench_b_tc.err
