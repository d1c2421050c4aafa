"""CUDA-graph replay against the step-by-step paths, BASELINE config 1 (B = 1, K = 100, T = 50) and config 2
(B = K = 4096, T = 100): the fused T-step filter (fused.GraphedFilter), and infer() on the torch-eager user
model captured whole (inference.GraphedInfer).  One JSON line per case."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import fused, inference  # noqa: E402
from tests.models import lgssm  # noqa: E402

dev = torch.device("cuda", 0)
torch.distributions.Distribution.set_default_validate_args(False)
for B, K, T in [(1, 100, 50), (64, 1024, 50), (4096, 4096, 100)]:
    ys = torch.from_numpy(lgssm.simulate(T, B, seed=1)).to(dev)
    model = fused.ScalarLinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 0.0, 0.5, device=dev)
    f = fused.GraphedFilter(model, T, B, K)
    eager = lgssm.bootstrap_filter(device=dev)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def stepwise():
        with torch.no_grad():
            return inference.infer("smc", ys, *model.callables(), K, return_log_marginal_likelihood=True, return_latents=False)

    def eager_path():
        with torch.no_grad():
            return inference.infer("smc", ys, *eager, K, return_log_marginal_likelihood=True, return_latents=False)

    obs_list = [ys[t] for t in range(T)]
    ge = inference.GraphedInfer("smc", obs_list, *eager, K, return_log_marginal_likelihood=True, return_latents=False)

    reps = 200 if B * K < 1e6 else 10
    tge = timed(lambda: ge(obs_list), max(3, reps // 4))
    tg = timed(lambda: f(ys, clone=False), reps)
    ts = timed(stepwise, max(3, reps // 10))
    te = timed(eager_path, max(3, reps // 20))
    print(json.dumps({"B": B, "K": K, "T": T, "graph_replay_ms": round(tg * 1e3, 3), "fused_stepwise_ms": round(ts * 1e3, 3),
                      "generic_eager_ms": round(te * 1e3, 3), "generic_eager_graph_replay_ms": round(tge * 1e3, 3),
                      "graph_particle_steps_per_s": B * K * T / tg}), flush=True)
