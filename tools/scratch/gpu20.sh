cd $GRAFT_REPO_ROOT
echo "baseline(in-tree)"; timeout 120 python tools/bench_field.py 2>&1 | grep '^{'
for v in lds lds_b5 ldsm ldsm_b4 ldsm_b5; do echo $v; NSIG_LIB=$PWD/tools/scratch/libs/libnsig_$v.so timeout 120 python tools/bench_field.py 2>&1 | grep '^{'; done
