"""Time infer() on the fused scalar linear-Gaussian model (the `e2e` leg of bench.py without the copies' bookkeeping):
B = K = 4096, T = 100, host observations in, host log-evidence out; ms per pass.  AESMC_B200_LIB selects the build."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _lib, fused, inference  # noqa: E402
from tests.models import lgssm  # noqa: E402

dev = torch.device("cuda", 0)
B = K = 4096
T = 100
obs_host = torch.from_numpy(lgssm.simulate(T, B, seed=7)).pin_memory()
out_host = torch.empty(B, dtype=torch.float32).pin_memory()
model = fused.ScalarLinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 0.0, 0.5, device=dev)
u = np.random.default_rng(0).random((T - 1, B))


def one_pass():
    dobs = obs_host.to(dev, non_blocking=True)
    with torch.no_grad():
        r = inference.infer("smc", [dobs[t] for t in range(T)], *model.callables(), K, return_log_marginal_likelihood=True,
                            return_latents=False, return_log_weight=False, uniforms=u)
    out_host.copy_(r["log_marginal_likelihood"], non_blocking=True)
    torch.cuda.current_stream().synchronize()


for _ in range(2):
    one_pass()
t0 = time.perf_counter()
for _ in range(5):
    one_pass()
ms = (time.perf_counter() - t0) * 1e3 / 5
print("%s: %.2f ms per pass, evidence mean %.6f" % (os.path.basename(_lib.LIB_PATH), ms, float(out_host.mean())), flush=True)
