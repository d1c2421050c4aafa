"""BASELINE configs 3 and 4 end to end through the public API, eager and as CUDA-graph replays.

  C3  10-D LGSSM with dense transition/emission and a learned Gaussian proposal (tests/models/lgssm_dense.py),
      B = K = 1024: infer('smc') for the log-evidence -- vector latents, D = 10 floats gathered per particle.
  C4  AESMC training of the nonlinear SSM with an MLP proposal (tests/models/nonlinear.py): get_loss('aesmc')
      forward + backward + Adam, one process (the data-parallel loop adds one flattened NCCL all-reduce per step).

One JSON line per case:   python scripts/bench_configs.py > profiles/r1_configs_3_4.jsonl
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import inference, losses, train  # noqa: E402
from tests.models import lgssm_dense, nonlinear  # noqa: E402

dev = torch.device("cuda", 0)
torch.distributions.Distribution.set_default_validate_args(False)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- config 3 ------------------------------------------------------------------------------------------
dx = dy = 10
B, K, T = 1024, 1024, 20
A, C = lgssm_dense.make_system(dx, dy, seed=3, device=dev)
ys = [y.to(dev) for y in lgssm_dense.simulate(A, C, T, B, 1.0, 0.5, 0.5, seed=4)]
init = lgssm_dense.Initial(dx, 1.0, dev)
trans, emis = lgssm_dense.Transition(A, 0.5).to(dev), lgssm_dense.Emission(C, 0.5).to(dev)
torch.manual_seed(0)
prop = lgssm_dense.Proposal(dx, dy).to(dev)


def eager3():
    with torch.no_grad():
        return inference.infer("smc", ys, init, trans, emis, prop, K, return_log_marginal_likelihood=True, return_latents=False)


g3 = inference.GraphedInfer("smc", ys, init, trans, emis, prop, K, return_log_marginal_likelihood=True, return_latents=False)
te, tg = timed(eager3, 5), timed(lambda: g3(ys), 10)
# the same four modules linked into the fused vector path (aesmc_lgv_propose_f32 + step kernel: 2 launches per time step)
from aesmc_b200 import fused  # noqa: E402
fused.link_dense(init, trans, emis, prop)
tf = timed(eager3, 10)
ys_host = torch.stack(ys).cpu().pin_memory()


def fused3_e2e():  # host observations in, log-evidence out
    with torch.no_grad():
        obs = ys_host.to(dev, non_blocking=True)
        r = inference.infer("smc", obs, init, trans, emis, prop, K, return_log_marginal_likelihood=True, return_latents=False)
        return r["log_marginal_likelihood"].cpu()


tfe = timed(fused3_e2e, 10)
del prop._aesmc_b200_fused
print(json.dumps({"config": 3, "model": "10-D dense LGSSM, learned Gaussian proposal", "B": B, "K": K, "T": T, "D": dx,
                  "infer_eager_ms": round(te * 1e3, 2), "infer_graph_replay_ms": round(tg * 1e3, 2),
                  "infer_fused_ms": round(tf * 1e3, 3), "infer_fused_e2e_host_obs_ms": round(tfe * 1e3, 3),
                  "particle_steps_per_s_eager": B * K * T / te, "particle_steps_per_s_graph": B * K * T / tg,
                  "particle_steps_per_s_fused": B * K * T / tf}), flush=True)

# ---- config 4 ------------------------------------------------------------------------------------------
for B, K, T in [(64, 1024, 20), (512, 1024, 20)]:
    torch.manual_seed(0)
    np.random.seed(0)
    init = nonlinear.Initial(dev)
    loader = train.get_synthetic_dataloader(init, nonlinear.Transition().to(dev), nonlinear.Emission().to(dev), T, B)
    batch = next(iter(loader))
    trans, emis, prop = nonlinear.Transition(scale=2.0).to(dev), nonlinear.Emission(mult=0.03).to(dev), nonlinear.Proposal().to(dev)
    opt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=1e-3)

    def eager4():
        opt.zero_grad()
        loss = losses.get_loss(batch, K, "aesmc", init, trans, emis, prop)
        loss.backward()
        opt.step()

    te = timed(eager4, 5)
    gopt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=1e-3, capturable=True)
    gstep = train.GraphedTrainStep(batch, K, "aesmc", init, trans, emis, prop, gopt)
    tg = timed(lambda: gstep(batch), 10)
    print(json.dumps({"config": 4, "model": "nonlinear SSM, MLP proposal (hidden 32), AESMC ELBO + Adam", "B": B, "K": K, "T": T,
                      "train_step_eager_ms": round(te * 1e3, 2), "train_step_graph_replay_ms": round(tg * 1e3, 2),
                      "particle_steps_per_s_eager": B * K * T / te, "particle_steps_per_s_graph": B * K * T / tg}), flush=True)
