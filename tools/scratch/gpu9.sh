cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_raymarching_gpu.py tests/test_ref_cuda_parity.py tests/test_render_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/profile_step.py --out gpurun_out/profile_step2.txt > /dev/null 2>&1; grep -E "k_march|k_field|k_composite|ms/step" gpurun_out/profile_step2.txt | cut -c1-110
timeout 600 python bench.py --no-cpu-baseline --no-render 2>/dev/null | tail -1 | cut -c1-200
