set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-400
for m in masks; do NSIG_BWD=$m timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_$m.log 2>&1; tail -n 1 gpurun_out/bench_field_$m.log; done
for c in 1 2 3 4; do NSIG_BWD=tc NSIG_TC_CTAS_PER_SM=$c timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_tc$c.log 2>&1; tail -n 1 gpurun_out/bench_field_tc$c.log; done
timeout 900 python bench.py > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_e.json; tail -3 gpurun_out/r02_bench_e.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>&1; tail -c 600 gpurun_out/r02_bench_ref.json
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
