set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
for v in base xpair; do
  NSIG_LIB=tools/scratch/libs/libnsig_$v.so timeout 300 python tools/bench_field.py --rays 8704 > gpurun_out/bench_field_$v.log 2>&1; tail -1 gpurun_out/bench_field_$v.log
done
for v in base xpair; do
  NSIG_LIB=tools/scratch/libs/libnsig_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_fwd' -c 2 -o gpurun_out/r02_fwd_$v python tools/bench_field.py --rays 8704 --iters 1 > gpurun_out/ncu_fwd_$v.log 2>&1; tail -2 gpurun_out/ncu_fwd_$v.log
done
