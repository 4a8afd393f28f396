#!/usr/bin/env python
"""Build an A/B variant of libnsig_b200.so with extra nvcc flags into tools/scratch/libs/ (select it with NSIG_LIB=...).

    python tools/build_variant.py NAME -DNSIG_BWD_MINB=5 [...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_signature_b200 import _build  # noqa: E402


def main():
    name, extra = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "tools", "scratch", "libs")
    obj_dir = os.path.join(out_dir, "obj_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    procs, objs = [], []
    for src in _build.SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_build._nvcc()] + _build.NVCC_FLAGS + extra + ["-Xptxas", "-v", "-c", os.path.join(_build.CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode:
            print(out)
            raise SystemExit(f"nvcc failed on {src}")
        if src == "field.cu":
            lines = out.splitlines()
            for i, l in enumerate(lines):
                if "Compiling entry function" in l and ("k_field_bwdILb0ELb0" in l or "k_field_fwdILb1ELb1" in l):
                    print(name, l.split("'")[1][:40], "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
    lib = os.path.join(out_dir, f"libnsig_{name}.so")
    subprocess.check_call([_build._nvcc(), "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
