cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(field|march|composite|msg)' -s 27 -c 9 -o gpurun_out/step_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-render --no-graph > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
