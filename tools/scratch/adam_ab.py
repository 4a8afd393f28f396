import hashlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nerf_signature_b200 import _lib
P = _lib.ptr
dev = torch.device("cuda:0")
md, log2_T = 32, 19
n = 2 << log2_T
gg = torch.Generator(device="cuda").manual_seed(11)
tabs = [torch.randn(n, device=dev, generator=gg) * 1e-2 for _ in range(2 * md)]
ms = [torch.randn(n, device=dev, generator=gg) * 1e-3 for _ in range(2 * md)]
vs = [torch.rand(n, device=dev, generator=gg) * 1e-6 for _ in range(2 * md)]
ptrs = torch.tensor([[t.data_ptr() for t in grp] for grp in (tabs, ms, vs)], dtype=torch.int64, device=dev)
steps = torch.arange(2 * md, dtype=torch.float32, device=dev) + 3.0
coef = torch.zeros(2 * md, 2, dtype=torch.float32, device=dev)
G = torch.randn(n, device=dev, generator=gg) * 65.536
msg = (torch.rand(md, device=dev, generator=gg) < 0.5).float()
scale = torch.tensor([65536.0], device=dev); finf = torch.zeros(1, device=dev)
for rng in ((0, 0), (4096, 131072)):
    _lib.call("nsig_msg_adam_step", P(ptrs), 2 * md, md, P(msg), P(G), P(steps), P(coef), P(scale), P(finf), 1e-2, 0.9, 0.99, 1e-15,
              log2_T, None, rng[0], rng[1], 0)
torch.cuda.synchronize()
h = hashlib.sha256()
for t in tabs + ms + vs:
    h.update(t.cpu().numpy().tobytes())
print("state sha256", h.hexdigest()[:16], "mode", os.environ.get("NSIG_ADAM_TMA", "default"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for i in range(20):
    flush.zero_()
    e0.record()
    _lib.call("nsig_msg_adam_step", P(ptrs), 2 * md, md, P(msg), P(G), P(steps), P(coef), P(scale), P(finf), 1e-2, 0.9, 0.99, 1e-15,
              log2_T, None, 0, 0, 0)
    e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
print("adam step %.1f us (incl. prepare kernel), %.2f TB/s" % (tot / 20 * 1e3, md * n * 4 * 6 / (tot / 20 * 1e-3) / 1e12))
