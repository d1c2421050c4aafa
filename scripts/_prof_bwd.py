import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _lib, _ops
dev = torch.device("cuda", 0)
B = K = 4096
gen = torch.Generator(device=dev).manual_seed(0)
lw = torch.randn(B, K, device=dev, generator=gen) - 1.4
u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
flags = _ops.new_flags(dev)
_, lse, idx, _ = _ops.smc_step(lw, None, None, u, None, flags, "exact", True)
for D in (1, 10):
    g = torch.randn(B, K, D, device=dev, generator=gen)
    gsrc = torch.empty_like(g)
    for _ in range(3):
        _lib.call("aesmc_gather_bwd_f32", _lib.ptr(g), _lib.ptr(idx), 0, B, K, D, _lib.ptr(gsrc), 1)
torch.cuda.synchronize()
