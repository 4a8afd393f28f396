set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_train_step_gpu.py -m gpu -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_f.log
tail -30 gpurun_out/pytest_f.log | cut -c1-300
for c in 1 2 3 4; do NSIG_BWD=tc NSIG_TC_CTAS_PER_SM=$c timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_tc$c.log 2>&1; tail -n 1 gpurun_out/bench_field_tc$c.log | cut -c1-330; done
for c in 2 4 6; do NSIG_BWD=tc_masks NSIG_TC_CTAS_PER_SM=$c timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_tcm$c.log 2>&1; tail -n 1 gpurun_out/bench_field_tcm$c.log | cut -c1-330; done
NSIG_BWD=tc_masks timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_bwd_tc_masks' -c 1 -o gpurun_out/r02_bwd_tcm python tools/bench_field.py --rays 8704 --iters 1 > gpurun_out/ncu_bwd_tcm.log 2>&1; tail -2 gpurun_out/ncu_bwd_tcm.log
NSIG_BWD=masks timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_bwd_masks|k_field_fwd' -c 2 -o gpurun_out/r02_fwd_bwd_masks python tools/bench_field.py --rays 8704 --iters 1 > gpurun_out/ncu_masks.log 2>&1; tail -2 gpurun_out/ncu_masks.log
NSIG_BWD=tc_masks timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_f_tcm.json 2> gpurun_out/r02_bench_f_tcm.err; tail -c 300 gpurun_out/r02_bench_f_tcm.json
