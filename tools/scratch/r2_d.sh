set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
NSIG_DEC_NO_PERSIST=1 timeout 300 python -m pytest tests/test_decoder_gpu.py -m gpu -q 2>&1 | tail -3
for m in recompute masks tc; do NSIG_BWD=$m timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_$m.log 2>&1; tail -n 1 gpurun_out/bench_field_$m.log; done
NSIG_BWD_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_bwd_tc' -c 1 -o gpurun_out/r02_bwd_tc python tools/bench_field.py --rays 8704 --iters 1 > gpurun_out/ncu_bwd_tc.log 2>&1; tail -2 gpurun_out/ncu_bwd_tc.log
NSIG_BWD=recompute timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_d_recompute.json 2> gpurun_out/r02_bench_d_recompute.err; tail -c 300 gpurun_out/r02_bench_d_recompute.json
timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_d_persist.json 2> gpurun_out/r02_bench_d_persist.err; tail -c 400 gpurun_out/r02_bench_d_persist.json; tail -3 gpurun_out/r02_bench_d_persist.err
NSIG_DEC_NO_PERSIST=1 timeout 600 python bench.py --no-extra --no-cpu-baseline --no-render > gpurun_out/r02_bench_d_nopersist.json 2> gpurun_out/r02_bench_d_nopersist.err; tail -c 400 gpurun_out/r02_bench_d_nopersist.json
timeout 300 python tools/bench_decoder.py > gpurun_out/bench_decoder_persist.log 2>&1; tail -n 3 gpurun_out/bench_decoder_persist.log
NSIG_DEC_NO_PERSIST=1 timeout 300 python tools/bench_decoder.py > gpurun_out/bench_decoder_nopersist.log 2>&1; tail -n 3 gpurun_out/bench_decoder_nopersist.log
