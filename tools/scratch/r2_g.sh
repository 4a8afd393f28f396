set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/profile_step.py --out gpurun_out/r02_step_breakdown.txt > gpurun_out/profile_step.log 2>&1; head -45 gpurun_out/r02_step_breakdown.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-extra --no-graph > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_field|k_march|k_composite|k_msg|k_grad|k_flat' -o gpurun_out/r02_step_main python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
