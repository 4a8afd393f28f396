"""CPU tests (-m "not gpu") of nerf_signature_b200/raymarching/backend.py - the drop-in for the reference's
`raymarching/backend.py` (the pybind11 call surface of raymarching/src/raymarching.h:7-18 over the C ABI): the ten names,
their parameter lists, and that every argument reaches the C entry point in the position include/nsig.h gives the parameter
of the same name.  No kernel runs: `_lib.call` is replaced by a recorder."""
import inspect
import os
import re

import pytest
import torch

from nerf_signature_b200 import _lib
from nerf_signature_b200.raymarching import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NSIG_REFERENCE", "/root/reference")
NAMES = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
         "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "composite_rays"]


def _c_params(header_text, fname):
    """[(type, name)] of a C declaration `... fname(type name, ...)`."""
    m = re.search(r"\b%s\s*\(([^)]*)\)" % re.escape(fname), header_text)
    assert m, fname
    out = []
    for prm in m.group(1).split(","):
        toks = prm.replace("*", " * ").split()
        out.append((" ".join(toks[:-1]), toks[-1]))
    return out


def _header():
    return re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nsig.h")).read(), flags=re.S)


def test_backend_has_the_ten_functions_with_the_c_abi_parameter_lists():
    h = _header()
    assert sorted(n for n in dir(backend._backend) if not n.startswith("_")) == sorted(NAMES)
    for name in NAMES:
        want = [p for _, p in _c_params(h, "nsig_" + name) if p not in ("scratch", "stream")]
        got = list(inspect.signature(getattr(backend._backend, name)).parameters)
        assert got == want, (name, got, want)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "raymarching", "src", "raymarching.h")),
                    reason="needs the reference sources (build container only)")
def test_backend_parameter_lists_are_the_reference_header_s():
    """raymarching/src/raymarching.h:7-18: same function names, same parameter names in the same order - which is what lets
    the reference's raymarching.py call `_backend.<fn>(...)` positionally without an edit."""
    ref = open(os.path.join(REF, "raymarching", "src", "raymarching.h")).read()
    for name in NAMES:
        want = [p for _, p in _c_params(ref, name)]
        got = list(inspect.signature(getattr(backend._backend, name)).parameters)
        assert got == want, (name, got, want)
    # and the reference wrapper calls nothing else on its backend
    used = set(re.findall(r"_backend\.(\w+)\(", open(os.path.join(REF, "raymarching", "raymarching.py")).read()))
    assert used == set(NAMES)


def test_every_argument_reaches_the_c_parameter_of_the_same_name(monkeypatch):
    h = _header()
    calls = []
    monkeypatch.setattr(_lib, "call", lambda fname, *args: calls.append((fname, args)))
    monkeypatch.setattr(backend, "_P", lambda t: None if t is None else ("ptr", id(t)))
    for name in NAMES:
        cparams = _c_params(h, "nsig_" + name)
        kwargs, scalars = {}, {}
        for i, (ctype, pname) in enumerate(cparams):
            if pname in ("scratch", "stream"):
                continue
            if "*" in ctype:
                kwargs[pname] = torch.zeros(4)                      # a distinct tensor object per pointer parameter
            else:
                scalars[pname] = kwargs[pname] = (i + 2) if "int" in ctype else (i + 0.5)
        calls.clear()
        getattr(backend._backend, name)(**kwargs)
        assert len(calls) == 1 and calls[0][0] == "nsig_" + name
        args = calls[0][1]
        expect_n = len([1 for _, p in cparams if p != "stream"])    # _lib.call appends the stream itself
        assert len(args) == expect_n, (name, len(args), expect_n)
        for (ctype, pname), a in zip(cparams, args):
            if pname == "scratch":
                assert a is not None
            elif "*" in ctype:
                assert a == ("ptr", id(kwargs[pname])), (name, pname)
            else:
                assert a == scalars[pname] and type(a) is (int if "int" in ctype else float), (name, pname, a)


def test_backend_has_no_cpu_path():
    with pytest.raises(_lib.NsigError):
        backend._backend.morton3D(torch.zeros(4, 3, dtype=torch.int32), 4, torch.zeros(4, dtype=torch.int32))
    with pytest.raises(_lib.NsigError):
        backend._backend.packbits(torch.zeros(4, 16)[:, ::2], 1, 0.5, torch.zeros(1, dtype=torch.uint8))   # strided view
