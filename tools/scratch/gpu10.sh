cd $GRAFT_REPO_ROOT
timeout 300 python tools/profile_step.py --out gpurun_out/profile_step4.txt > /dev/null 2>&1; head -40 gpurun_out/profile_step4.txt | cut -c1-150
