cd $GRAFT_REPO_ROOT
for v in "-DNSIG_FWD_MINB=4" "-DNSIG_FWD_MINB=5" "-DNSIG_FWD_MINB=6"; do
NSIG_NVCC_EXTRA="$v" python -c "
import sys; sys.path.insert(0,'.')
from nerf_signature_b200 import _build; _build.build_library(force=True)" >/dev/null 2>&1
echo "variant [$v]"; python tools/bench_field.py 2>&1 | tail -1
done
