cd $GRAFT_REPO_ROOT
echo "== default"; timeout 200 python tools/scratch/adam_cmp2.py 2>&1 | grep -v Warn | tail -22
echo "== NSIG_NO_DIRECT_SINK=1"; NSIG_NO_DIRECT_SINK=1 timeout 200 python tools/scratch/adam_cmp2.py 2>&1 | grep -v Warn | tail -22
