"""CPU dry run of tests/test_zz_backend_shim_gpu.py: `.cuda()` becomes the identity, device allocations land on the host
and `_lib.call` dispatches to the C oracle (whose entry points have the argument order of raymarching.h), so the test's own
plumbing - shapes, dtypes, caller-allocated buffers, in-place alive-ray state - is exercised without a GPU.
Usage: python tools/dryrun_backend_shim.py"""
import sys, ctypes as c, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from oracle import cpu as oc
oc.build()
olib = oc.lib()
from nerf_signature_b200 import _lib
from nerf_signature_b200.raymarching import backend
# --- emulate the device: .cuda() is identity, device="cuda" allocations land on the CPU
torch.Tensor.cuda = lambda self,*a,**k: self
for fn in ("empty","zeros"):
    orig=getattr(torch,fn)
    def mk(orig):
        def f(*a,**k):
            k.pop("device",None); return orig(*a,**k)
        return f
    setattr(torch,fn,mk(orig))
backend._P = lambda t: None if t is None else c.c_void_p(t.data_ptr())
def fake_call(name,*args):
    fn=getattr(olib,"oracle_"+name[len("nsig_"):])
    sig=_lib._SIGNATURES[name][0]
    conv=[]
    args=list(args)
    if name=="nsig_march_rays_train": args=args[:-1]; sig=sig[:-2]   # no scratch in the oracle
    else: sig=sig[:-1]
    assert len(args)==len(sig),(name,len(args),len(sig))
    for a,t in zip(args,sig):
        conv.append(a if t is c.c_void_p else t(a))
    fn.restype=None
    fn(*conv)
_lib.call=fake_call
backend._lib.call=fake_call
import test_zz_backend_shim_gpu as T
T.test_utilities_through_the_backend(oc); print("utilities ok")
T.test_training_march_and_composite_through_the_backend(oc); print("train ok")
T.test_inference_march_and_composite_through_the_backend(oc); print("inference ok")
