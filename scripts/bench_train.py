"""get_loss('aesmc') forward + backward throughput with the reference-style LGSSM user model
(tests/models/lgssm.py: learnable transition/emission multipliers, two-Linear proposal), torch-eager model
ops + the step kernel with its autograd backward, eager and replayed as one CUDA graph.  SURVEY 6 measured the reference on CPU at the same
shapes (fwd+bwd 7.3e6 particle-steps/s at B=64, K=4096, T=10 on 8 vCPU)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import losses, train  # noqa: E402
from tests.models import lgssm  # noqa: E402

dev = torch.device("cuda", 0)
torch.distributions.Distribution.set_default_validate_args(False)
for B, K, T in [(64, 4096, 10), (256, 4096, 20), (1024, 4096, 50)]:
    torch.manual_seed(0)
    init = lgssm.Initial(0.0, 1.0)
    trans, emis, prop = lgssm.Transition(0.9, 1.0).to(dev), lgssm.Emission(1.0, 0.5).to(dev), lgssm.Proposal(0.8, 0.7).to(dev)
    obs = [torch.randn(B, device=dev) for _ in range(T)]
    opt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=1e-3)

    def step():
        opt.zero_grad()
        loss = losses.get_loss(obs, K, "aesmc", init, trans, emis, prop)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    # the same step captured once as a CUDA graph (train.GraphedTrainStep) and replayed
    gopt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=1e-3, capturable=True)
    gstep = train.GraphedTrainStep(obs, K, "aesmc", init, trans, emis, prop, gopt)
    for _ in range(3):
        gstep(obs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        gstep(obs)
    torch.cuda.synchronize()
    dg = (time.perf_counter() - t0) / reps
    print(json.dumps({"B": B, "K": K, "T": T, "fwd_bwd_step_ms": round(dt * 1e3, 2), "particle_steps_per_s": B * K * T / dt,
                      "graph_replay_step_ms": round(dg * 1e3, 2), "graph_particle_steps_per_s": B * K * T / dg,
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}), flush=True)
