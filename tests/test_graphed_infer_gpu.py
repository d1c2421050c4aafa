"""inference.GraphedInfer: infer() on an arbitrary torch user model captured as one CUDA graph."""
import numpy as np
import pytest
import torch

from aesmc_b200 import inference
from oracle import kalman
from tests.models import lgssm

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def no_distribution_validation():
    # argument validation synchronises with the host (.all()), which a graph capture forbids
    old = torch.distributions.Distribution._validate_args
    torch.distributions.Distribution.set_default_validate_args(False)
    yield
    torch.distributions.Distribution.set_default_validate_args(old)


def test_graph_replay_tracks_kalman_and_follows_new_observations(cuda):
    T, B, K = 30, 8, 2048
    ys = lgssm.simulate(T, B, seed=3)
    ys2 = lgssm.simulate(T, B, seed=4)
    obs = [torch.from_numpy(y).to(cuda) for y in ys]
    torch.manual_seed(0)
    g = inference.GraphedInfer("smc", obs, *lgssm.bootstrap_filter(device=cuda), K, return_log_marginal_likelihood=True,
                               return_latents=False)
    for data in (ys, ys2, ys):
        res = g([torch.from_numpy(y).to(cuda) for y in data])
        g.check()
        exact = kalman.lgssm1d_log_evidence(data, 0.0, 1.0, 0.9, 1.0, 1.0, 0.25)
        err = np.abs(res["log_marginal_likelihood"].cpu().numpy() - exact)
        assert err.max() < 0.6, err
        assert res["latents"] is None and res["log_weight"].shape == (B, K) and res["last_latent"].shape == (B, K)


def test_seed_controls_a_replay_and_replays_differ(cuda):
    T, B, K = 12, 4, 256
    obs = [torch.randn(B, device=cuda) for _ in range(T)]
    torch.manual_seed(0)
    init, trans, emis, prop = lgssm.Initial(0.0, 1.0), lgssm.Transition(0.9, 1.0).to(cuda), lgssm.Emission(1.0, 0.5).to(cuda), \
        lgssm.Proposal(0.8, 0.7).to(cuda)
    g = inference.GraphedInfer("smc", obs, init, trans, emis, prop, K, return_log_marginal_likelihood=True,
                               return_log_weights=True, return_ancestral_indices=True)
    torch.manual_seed(5)
    a = {k: (torch.stack(v) if isinstance(v, list) else v).clone() for k, v in g().items() if v is not None and k != "latents"}
    b_lml = g()["log_marginal_likelihood"].clone()          # the generator moved on: a different draw
    torch.manual_seed(5)
    c = g()
    assert not torch.equal(a["log_marginal_likelihood"], b_lml)
    assert torch.equal(a["log_marginal_likelihood"], c["log_marginal_likelihood"])
    assert torch.equal(a["ancestral_indices"], torch.stack(c["ancestral_indices"]))
    assert len(c["latents"]) == T and c["ancestral_indices"][0].dtype == torch.int64


def test_importance_sampling_and_argument_checks(cuda):
    T, B, K = 6, 3, 512
    obs = [torch.randn(B, device=cuda) for _ in range(T)]
    g = inference.GraphedInfer("is", obs, *lgssm.bootstrap_filter(device=cuda), K, return_log_marginal_likelihood=True)
    res = g()
    assert torch.isfinite(res["log_marginal_likelihood"]).all() and len(res["latents"]) == T
    with pytest.raises(ValueError):
        g(obs[:-1])
    with pytest.raises(ValueError):
        inference.GraphedInfer("smc", obs, *lgssm.bootstrap_filter(device=cuda), K, uniforms=None)
    with pytest.raises(ValueError):
        inference.GraphedInfer("smc", [o.cpu() for o in obs], *lgssm.bootstrap_filter(device=cuda), K)


def test_graphed_train_step_learns(cuda):
    """train.GraphedTrainStep: forward + backward + Adam replayed as one graph moves the parameters the way
    the eager loop does (the loss of the LGSSM proposal falls)."""
    from aesmc_b200 import losses, train
    T, B, K = 8, 32, 256
    torch.manual_seed(0)
    init = lgssm.Initial(0.0, 1.0)
    trans, emis, prop = lgssm.Transition(0.5, 1.0).to(cuda), lgssm.Emission(1.0, 0.5).to(cuda), lgssm.Proposal(0.1, 0.1).to(cuda)
    ys = lgssm.simulate(T, B, seed=2)
    obs = [torch.from_numpy(y).to(cuda) for y in ys]
    params = list(train.get_chained_params(trans, emis, prop))
    before = [p.detach().clone() for p in params]
    with pytest.raises(ValueError):
        train.GraphedTrainStep(obs, K, "aesmc", init, trans, emis, prop, torch.optim.Adam(params, lr=1e-2))
    opt = torch.optim.Adam(params, lr=2e-2, capturable=True)
    step = train.GraphedTrainStep(obs, K, "aesmc", init, trans, emis, prop, opt)
    first = float(step())
    for _ in range(150):
        loss = step(obs)
    last = float(loss)
    assert np.isfinite(first) and np.isfinite(last) and last < first - 0.05, (first, last)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, params))
    # the trained model evaluated eagerly agrees with what the graph reports (same objective, fresh noise)
    with torch.no_grad():
        eager = float(losses.get_loss(obs, K, "aesmc", init, trans, emis, prop))
    assert abs(eager - last) < 0.5, (eager, last)


def test_graphed_train_step_nonlinear_mlp_proposal(cuda):
    """BASELINE config 4's model family (nonlinear SSM, MLP proposal with BATCH_EXPANDED / FULLY_EXPANDED
    distributions) under the captured training step, both objectives."""
    from aesmc_b200 import train
    from tests.models import nonlinear
    torch.manual_seed(0)
    np.random.seed(0)
    init = nonlinear.Initial(cuda)
    true_trans, true_emis = nonlinear.Transition().to(cuda), nonlinear.Emission().to(cuda)
    loader = train.get_synthetic_dataloader(init, true_trans, true_emis, num_timesteps=10, batch_size=32)
    batch = next(iter(loader))
    for algorithm in ("aesmc", "iwae"):
        trans, emis = nonlinear.Transition(scale=2.0).to(cuda), nonlinear.Emission(mult=0.03).to(cuda)
        prop = nonlinear.Proposal().to(cuda)
        opt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=5e-3, capturable=True)
        step = train.GraphedTrainStep(batch, 256, algorithm, init, trans, emis, prop, opt)
        seen = [float(step(batch)) for _ in range(60)]
        assert all(np.isfinite(seen)), seen[:5]
        assert np.mean(seen[-5:]) < np.mean(seen[:5]), (algorithm, np.mean(seen[:5]), np.mean(seen[-5:]))


def test_graphed_prior_sampler_and_dataset(cuda):
    """train.GraphedPriorSampler: statistics.sample_from_prior as one CUDA graph -- fresh draws per replay with the
    eager sampler's distribution, feeding the captured training step through SyntheticDataset(graphed=True)."""
    from aesmc_b200 import statistics, train
    from tests.models import nonlinear
    torch.manual_seed(0)
    init = nonlinear.Initial(cuda)
    true_trans, true_emis = nonlinear.Transition().to(cuda), nonlinear.Emission().to(cuda)
    T, B = 8, 4096
    launches = None
    sampler = train.GraphedPriorSampler(init, true_trans, true_emis, T, B, keep_latents=True)
    a = [o.clone() for o in sampler()]
    b = sampler(clone=True)
    assert len(a) == T and a[0].shape == b[0].shape and a[0].is_cuda
    assert not torch.equal(a[0], b[0]) and not torch.equal(a[-1], b[-1])          # fresh noise every replay
    with torch.no_grad():
        _, eager = statistics.sample_from_prior(init, true_trans, true_emis, T, B)
    for t in (0, T - 1):                                                          # same distribution as the eager sampler
        ge, ee = b[t].double().flatten(), eager[t].double().flatten()
        se = (ge.var() / B + ee.var() / B).sqrt().item()
        assert abs(ge.mean().item() - ee.mean().item()) < 6 * se, (t, ge.mean().item(), ee.mean().item())
        assert 0.8 < (ge.std() / ee.std()).item() < 1.25
    torch.manual_seed(7)
    s1 = train.GraphedPriorSampler(init, true_trans, true_emis, 3, 16)
    x1 = s1(clone=True)
    torch.manual_seed(7)
    s2 = train.GraphedPriorSampler(init, true_trans, true_emis, 3, 16)
    assert all(torch.equal(p, q) for p, q in zip(x1, s2(clone=True)))               # torch.manual_seed controls the stream
    # the dataset / dataloader front end, feeding a captured training step
    loader = train.get_synthetic_dataloader(init, true_trans, true_emis, num_timesteps=6, batch_size=32, graphed=True)
    it = iter(loader)
    batch = next(it)
    assert len(batch) == 6 and batch[0].shape[0] == 32 and batch[0].is_cuda
    trans, emis, prop = nonlinear.Transition(scale=2.0).to(cuda), nonlinear.Emission(mult=0.03).to(cuda), nonlinear.Proposal().to(cuda)
    opt = torch.optim.Adam(train.get_chained_params(trans, emis, prop), lr=5e-3, capturable=True)
    step = train.GraphedTrainStep(batch, 128, "aesmc", init, trans, emis, prop, opt)
    seen = [float(step(next(it))) for _ in range(40)]
    assert all(np.isfinite(seen)) and np.mean(seen[-8:]) < np.mean(seen[:8])
