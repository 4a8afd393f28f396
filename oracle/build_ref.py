"""Build the UNMODIFIED reference `_raymarching` CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The sources are compiled from where they lie under
/root/reference (raymarching/src/raymarching.cu + bindings.cpp); nothing is copied
into this repository.  The only deviation from the reference's own recipe
(raymarching/backend.py:6-9) is `-std=c++17` instead of `-std=c++14`, because
torch 2.11 headers need C++17, plus an explicit sm_100a gencode so
the cubin runs on B200.  Default nvcc floating-point flags are kept (-fmad=true,
IEEE div/sqrt) because integer outputs of the march kernel depend on them.

The resulting oracle/_ref/_raymarching.so is git-ignored but travels to the GPU
box with the gpurun snapshot, where `tests/` use it as the bit-exactness oracle
for the march/composite kernels (it cannot execute here: no GPU).
"""
import os
import sys

REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(verbose=True):
    src = os.path.join(REF, "raymarching", "src")
    if not os.path.isdir(src):
        print(f"[oracle/_ref] reference not present at {REF}; skipping")
        return None
    so = os.path.join(OUT, "_raymarching.so")
    if os.path.exists(so):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    nvcc_flags = ['-O3', '-std=c++17',
                  '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
                  '-U__CUDA_NO_HALF2_OPERATORS__']
    load(name='_raymarching', extra_cflags=['-O3', '-std=c++17'],
         extra_cuda_cflags=nvcc_flags,
         sources=[os.path.join(src, f) for f in ('raymarching.cu', 'bindings.cpp')],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so


if __name__ == "__main__":
    print(build())
