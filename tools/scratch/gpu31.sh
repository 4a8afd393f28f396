cd $GRAFT_REPO_ROOT
for v in base r3 r4; do
 if [ $v = base ]; then unset NSIG_LIB; else export NSIG_LIB=$PWD/tools/scratch/libs/libnsig_$v.so; fi
 timeout 200 python tools/bench_render.py --views 5 2>&1 | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', {k:(round(v['fused_whole_frame']['ms_per_frame'],2), round(v['fused_staged_4096']['ms_per_frame'],2)) for k,v in d.items()})"
done
