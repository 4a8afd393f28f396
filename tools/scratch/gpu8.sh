cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-graph > gpurun_out/launches_v2_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nsig' -s 24 -c 14 -o gpurun_out/step_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-render --no-graph > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
timeout 600 python bench.py > gpurun_out/bench_v3.log 2>&1; tail -1 gpurun_out/bench_v3.log | cut -c1-400
