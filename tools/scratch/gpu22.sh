cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^k_ -f -o gpurun_out/step_v5 python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log; ls -la gpurun_out/*.ncu-rep
