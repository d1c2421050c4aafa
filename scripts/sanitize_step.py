"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck) over every step-kernel variant:
register-blocked (regular and irregular K, D = 1 and D > 1), generic, multi-CTA, fused, both modes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _ops, fused  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
for mode in ("exact", "fast"):
    for B, K, D in [(5, 4096, 1), (3, 1000, 1), (3, 2048, 3), (2, 1023, 1), (2, 30000, 1), (1, 70000, 2)]:
        a, b, c = [torch.randn(B, K, device=dev, generator=gen) for _ in range(3)]
        x = torch.randn(B, K, D, device=dev, generator=gen) if D > 1 else torch.randn(B, K, device=dev, generator=gen)
        u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
        flags = _ops.new_flags(dev)
        out = _ops.smc_step(a, b, c, u, x, flags, mode, True)
        out2 = _ops.smc_step(a, b, c, None, None, flags, mode, False)
        torch.cuda.synchronize()
        assert int(flags.item()) == 0 and int(out[2].max()) < K
    model = fused.ScalarLinearGaussianSSM(device=dev)
    obs = torch.randn(4, 3, device=dev, generator=gen)
    with torch.no_grad():
        r = fused.infer_fused(model, obs, 512, return_log_marginal_likelihood=True, resampling_mode=mode)
    torch.cuda.synchronize()
    assert torch.isfinite(r["log_marginal_likelihood"]).all()
print("sanitize driver ok")
