set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
for v in 0 1; do timeout 60 tools/scratch/tc_probe $v > gpurun_out/tc_probe_$v.log 2>&1; echo "probe $v rc=$?"; cat gpurun_out/tc_probe_$v.log; done
NSIG_BWD_TC=0 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_field_gpu.py::test_tcgen05_backward_matches_mma_sync_backward > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log | cut -c1-300
NSIG_BWD_TC=0 timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_mma.log 2>&1; tail -n 1 gpurun_out/bench_field_mma.log
timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_tc.log 2>&1; tail -n 1 gpurun_out/bench_field_tc.log
