cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_fwd' -s 3 -c 2 -o gpurun_out/field_v2 python tools/bench_field.py --iters 3 > gpurun_out/ncu_v2.log 2>&1
NSIG_NVCC_EXTRA="-DNSIG_GATHER_V1" python -c "
import sys; sys.path.insert(0,'.')
from nerf_signature_b200 import _build; _build.build_library(force=True)" >/dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_fwd' -s 3 -c 2 -o gpurun_out/field_v1h python tools/bench_field.py --iters 3 > gpurun_out/ncu_v1h.log 2>&1
ls -la gpurun_out
