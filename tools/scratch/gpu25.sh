cd $GRAFT_REPO_ROOT
NSIG_LIB=$PWD/tools/scratch/libs/libnsig_pdl0.so timeout 120 python tools/bench_decoder.py 2>&1 | grep "^decoder"
