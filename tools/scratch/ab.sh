set -e
cd $GRAFT_REPO_ROOT
for v in "-DNSIG_NO_PAIR_GATHER" ""; do
  NSIG_NVCC_EXTRA="$v" python -c "
import sys; sys.path.insert(0,'.')
from nerf_signature_b200 import _build; _build.build_library(force=True)" >/dev/null 2>&1
  echo "variant [$v]"; python tools/bench_field.py 2>&1 | tail -1
done
python -m pytest tests/test_field_gpu.py tests/test_hash_gpu.py tests/test_render_gpu.py -x -q 2>&1 | tail -3
