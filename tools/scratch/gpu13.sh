cd $GRAFT_REPO_ROOT
echo "== default"; timeout 200 python tools/scratch/adam_cmp.py 2>&1 | grep -v Warn | tail -7
echo "== no side streams at all"; NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 200 python tools/scratch/adam_cmp.py 2>&1 | grep -v Warn | tail -7
echo "== G16"; NSIG_DEC_WGRAD_G=16 NSIG_NO_SIDE_STREAMS=1 NSIG_DEC_NO_SIDE=1 timeout 200 python tools/scratch/adam_cmp.py 2>&1 | grep -v Warn | tail -7
for g in 8 12 16; do NSIG_DEC_WGRAD_G=$g timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder forward+"; done
