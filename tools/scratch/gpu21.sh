cd $GRAFT_REPO_ROOT
for v in mt2 mt2_b3 mt2_lds_b3; do echo $v; NSIG_LIB=$PWD/tools/scratch/libs/libnsig_$v.so timeout 120 python tools/bench_field.py 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['field_fwd_ms'], d['field_bwd_ms'])"; done
NSIG_LIB=$PWD/tools/scratch/libs/libnsig_mt2.so timeout 300 python -m pytest tests/test_field_gpu.py -x -q 2>&1 | tail -3
