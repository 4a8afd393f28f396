cd $GRAFT_REPO_ROOT
for v in base ldsm ldsm_b4 mt2 mt2_ldsm_b3; do
 if [ $v = base ]; then unset NSIG_LIB; else export NSIG_LIB=$PWD/tools/scratch/libs/libnsig_$v.so; fi
 timeout 120 python tools/bench_field.py 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v fwd/bwd ms', round(d['field_fwd_ms'],4), round(d['field_bwd_ms'],4))"
done
