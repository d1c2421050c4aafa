"""BASELINE configs 3 and 4 through the public API on the GPU: dense 10-D LGSSM (D = 10 ancestral
gather) against the matrix Kalman filter, and AESMC training of an MLP proposal on the nonlinear SSM
(single process; the 2-rank NCCL variant is tests/test_multi_gpu.py)."""
import numpy as np
import pytest
import torch

import aesmc_b200
from aesmc_b200 import inference, losses, train
from oracle import kalman
from tests.models import lgssm_dense, nonlinear

pytestmark = pytest.mark.gpu


def _dense_setup(dev, r, seed=1):
    dx = dy = 10
    s0, q = 1.0, 0.5
    A, C = lgssm_dense.make_system(dx, dy, seed=seed, device=dev)
    init = lgssm_dense.Initial(dx, s0, dev)
    trans = lgssm_dense.Transition(A, q)
    emis = lgssm_dense.Emission(C, r)
    return dx, dy, s0, q, A, C, init, trans, emis, lgssm_dense.PriorProposal(init, trans)


def test_dense_lgssm_matches_port_on_shared_noise(cuda):
    """D = 10 latents: same noise and uniforms through the CPU oracle port and the GPU path."""
    from tests.test_infer_gpu import fixed_noise
    from oracle import reference_port as port
    T, B, K = 6, 4, 1024
    u = np.random.default_rng(0).random((T - 1, B))
    out = {}
    for dev in ("cpu", cuda):
        dx, dy, s0, q, A, C, init, trans, emis, prop = _dense_setup(dev, r=1.0)
        ys = lgssm_dense.simulate(A, C, T, B, s0, q, 1.0, seed=2)
        fn = port.infer if dev == "cpu" else inference.infer
        with fixed_noise(5), torch.no_grad():
            out[dev] = fn("smc", [y.to(dev) for y in ys], init, trans, emis, prop, K, return_log_marginal_likelihood=True,
                          return_latents=True, return_ancestral_indices=True, uniforms=u)
    anc_ref = torch.stack(out["cpu"]["ancestral_indices"]).numpy()
    anc_got = torch.stack(out[cuda]["ancestral_indices"]).cpu().numpy()
    frac = (anc_ref != anc_got).mean()
    print("10-D LGSSM: index mismatch fraction vs CPU port under shared noise:", frac)
    assert frac < 2e-3
    np.testing.assert_allclose(out[cuda]["log_marginal_likelihood"].cpu().numpy(),
                               out["cpu"]["log_marginal_likelihood"].numpy(), rtol=1e-4, atol=1e-3)
    assert out[cuda]["latents"][0].shape == (B, K, 10)


def test_dense_lgssm_evidence_tracks_kalman(cuda):
    T, B, K = 10, 6, 16384
    r = 2.0   # weakly informative observations keep the 10-D bootstrap filter's variance moderate
    dx, dy, s0, q, A, C, init, trans, emis, prop = _dense_setup(cuda, r)
    ys = lgssm_dense.simulate(A, C, T, B, s0, q, r, seed=2)
    exact = kalman.lgssm_log_evidence(ys.numpy(), np.zeros(dx), s0 ** 2 * np.eye(dx), A.cpu().numpy(), q ** 2 * np.eye(dx),
                                      C.cpu().numpy(), r ** 2 * np.eye(dy))
    torch.manual_seed(0)
    np.random.seed(0)
    with torch.no_grad():
        res = inference.infer("smc", [y.to(cuda) for y in ys], init, trans, emis, prop, K,
                              return_log_marginal_likelihood=True, return_latents=False)
    err = np.abs(res["log_marginal_likelihood"].cpu().numpy() - exact)
    print("10-D LGSSM bootstrap filter: max |log Z_hat - log Z| =", err.max(), "at K =", K)
    assert err.max() < 1.0


def test_dense_lgssm_learned_proposal_trains(cuda):
    dx = dy = 10
    T, B, K = 8, 16, 256
    A, C = lgssm_dense.make_system(dx, dy, seed=3, device=cuda)
    ys = [y.to(cuda) for y in lgssm_dense.simulate(A, C, T, B, 1.0, 0.5, 0.5, seed=4)]
    init = lgssm_dense.Initial(dx, 1.0, cuda)
    trans = lgssm_dense.Transition(A, 0.5, learn=True).to(cuda)
    emis = lgssm_dense.Emission(C, 0.5, learn=True).to(cuda)
    torch.manual_seed(0)
    prop = lgssm_dense.Proposal(dx, dy).to(cuda)
    opt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=2e-2)
    history = []
    for it in range(40):
        opt.zero_grad()
        loss = losses.get_loss(ys, K, "aesmc", init, trans, emis, prop)
        loss.backward()
        opt.step()
        history.append(loss.item())
    assert all(np.isfinite(history)) and np.mean(history[-5:]) < np.mean(history[:5])
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in prop.parameters())


def test_nonlinear_ssm_training_loop(cuda):
    torch.manual_seed(0)
    np.random.seed(0)
    init = nonlinear.Initial(cuda)
    true_trans, true_emis = nonlinear.Transition().to(cuda), nonlinear.Emission().to(cuda)
    loader = train.get_synthetic_dataloader(init, true_trans, true_emis, num_timesteps=10, batch_size=32)
    fixed = [next(iter(loader))] * 40          # one fixed batch of sequences: the loss must go down on it
    trans, emis = nonlinear.Transition(scale=2.0).to(cuda), nonlinear.Emission(mult=0.03).to(cuda)
    prop = nonlinear.Proposal().to(cuda)
    seen = []
    train.train(fixed, 256, "aesmc", init, trans, emis, prop, num_epochs=1, optimizer_kwargs={"lr": 5e-3},
                callback=lambda e, i, loss, *models: seen.append(loss.item()))
    assert len(seen) == 40 and all(np.isfinite(seen))
    print("nonlinear SSM AESMC loss: first 5 %.2f -> last 5 %.2f" % (np.mean(seen[:5]), np.mean(seen[-5:])))
    assert np.mean(seen[-5:]) < np.mean(seen[:5])
    # the infinite synthetic loader and 'iwae' run through the same loop
    train.train(loader, 16, "iwae", init, trans, emis, prop, num_epochs=1, num_iterations_per_epoch=2)
    train.train(loader, 16, "aesmc", init, trans, emis, prop, num_epochs=2, num_iterations_per_epoch=1)
