"""Fused-model fast path (aesmc_b200.fused): the step kernel evaluates the scalar linear-Gaussian model
itself.  With injected noise it must reproduce the generic path (torch-eager user model + step kernel)
bit for bit; with its own Philox stream it must track the exact Kalman evidence."""
import numpy as np
import pytest
import torch

import aesmc_b200
from aesmc_b200 import fused, inference
from oracle import kalman
from tests.models import lgssm

pytestmark = pytest.mark.gpu

AFFINE = dict(p0_y=0.7, p0_off=0.1, p0_scale=0.8, pt_x=0.5, pt_y=0.4, pt_off=-0.05, pt_scale=0.6)


def eager_with_noise(model, obs, K, noise, u, **kw):
    """Generic path, every Normal.rsample fed from `noise` (t-th draw = noise[t], transposed when the
    distribution is batch-expanded and therefore sampled as [K, B])."""
    import torch.distributions.normal as tdn
    it = iter(noise)
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        z = next(it)
        return z if tuple(shape) == tuple(z.shape) else z.t().contiguous()

    tdn._standard_normal = fake
    try:
        with torch.no_grad():
            # the lambdas hide the bound methods, so infer() takes the generic path
            return inference.infer("smc", obs, lambda: model.initial(), lambda **k: model.transition(**k),
                                   lambda **k: model.emission(**k), lambda **k: model.proposal(**k), K, uniforms=u, **kw)
    finally:
        tdn._standard_normal = orig


@pytest.mark.parametrize("proposal", ["bootstrap", AFFINE])
@pytest.mark.parametrize("K", [256, 4096])
def test_fused_equals_generic_path_bitwise(cuda, proposal, K):
    T, B = 7, 5
    model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device=cuda)
    obs = torch.from_numpy(lgssm.simulate(T, B, seed=1)).to(cuda)
    gen = torch.Generator(device=cuda).manual_seed(0)
    noise = torch.randn(T, B, K, device=cuda, generator=gen)
    u = np.random.default_rng(0).random((T - 1, B))
    kw = dict(return_log_marginal_likelihood=True, return_latents=True, return_original_latents=True,
              return_log_weight=True, return_log_weights=True, return_ancestral_indices=True)
    ref = eager_with_noise(model, obs, K, noise, u, **kw)
    with torch.no_grad():
        got = fused.infer_fused(model, obs, K, uniforms=u, noise=noise, **kw)
    for t in range(T):
        assert torch.equal(got["original_latents"][t], ref["original_latents"][t]), ("latent", t)
        assert torch.equal(got["log_weights"][t], ref["log_weights"][t]), ("log_w", t)
        assert torch.equal(got["latents"][t], ref["latents"][t])
    for a, b in zip(got["ancestral_indices"], ref["ancestral_indices"]):
        assert torch.equal(a, b)
    assert torch.equal(got["log_marginal_likelihood"], ref["log_marginal_likelihood"])
    assert torch.equal(got["log_weight"], ref["log_weight"]) and torch.equal(got["last_latent"], ref["last_latent"])


def test_infer_dispatches_to_fused_and_tracks_kalman(cuda):
    T, B, K = 50, 16, 4096
    ys = lgssm.simulate(T, B, seed=9)
    exact = kalman.lgssm1d_log_evidence(ys, 0.0, 1.0, 0.9, 1.0, 1.0, 0.25)
    model = fused.ScalarLinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 0.0, 0.5, device=cuda)
    obs = torch.from_numpy(ys).to(cuda)
    assert fused.applicable(fused.model_of(*model.callables()), obs, K)
    launches = aesmc_b200._lib.launch_count()
    torch.manual_seed(0)
    np.random.seed(0)
    with torch.no_grad():
        res = inference.infer("smc", obs, *model.callables(), K, return_log_marginal_likelihood=True, return_latents=False)
    assert aesmc_b200._lib.launch_count() - launches == T        # one launch per time step, nothing else
    err = np.abs(res["log_marginal_likelihood"].cpu().numpy() - exact)
    print("fused bootstrap filter: max |log Z_hat - log Z| =", err.max())
    assert err.max() < 1.0 and err.mean() < 0.35      # O(sqrt(T/K)) per row (0.64 max at K = 1000)
    assert res["log_weight"].shape == (B, K) and res["latents"] is None
    # different torch seed -> different Philox stream; same seed -> identical result
    torch.manual_seed(1); np.random.seed(0)
    with torch.no_grad():
        r1 = inference.infer("smc", obs, *model.callables(), K, return_log_marginal_likelihood=True, return_latents=False)
    torch.manual_seed(1); np.random.seed(0)
    with torch.no_grad():
        r2 = inference.infer("smc", obs, *model.callables(), K, return_log_marginal_likelihood=True, return_latents=False)
    assert torch.equal(r1["log_marginal_likelihood"], r2["log_marginal_likelihood"])
    assert not torch.equal(r1["log_marginal_likelihood"], res["log_marginal_likelihood"])


def test_philox_normals_are_standard(cuda):
    model = fused.ScalarLinearGaussianSSM(0.0, 1.0, device=cuda)
    obs = torch.zeros(1, 64, device=cuda)
    torch.manual_seed(3)
    with torch.no_grad():
        x = fused.infer_fused(model, obs, 16384, return_latents=False)["last_latent"].double().flatten()
    n = x.numel()
    assert abs(x.mean().item()) < 5 / np.sqrt(n) and abs(x.var().item() - 1) < 5 * np.sqrt(2 / n)
    assert abs((x ** 4).mean().item() - 3) < 0.05 and abs((x ** 3).mean().item()) < 0.02
    assert x.unique().numel() > 0.99 * n


def test_fused_model_runs_through_generic_path_and_reference_conventions(cuda):
    # not applicable (CPU observations / odd K): the same callables take the generic path
    model = fused.ScalarLinearGaussianSSM(device=cuda)
    obs = [torch.randn(3, device=cuda) for _ in range(4)]
    res = inference.infer("smc", obs, *model.callables(), 10)
    assert res["log_weight"].shape == (3, 10)
    assert fused.model_of(model.initial, model.transition, model.emission, lambda **k: None) is None


@pytest.mark.parametrize("proposal", ["bootstrap", AFFINE])
def test_graphed_filter_equals_stepwise_path(cuda, proposal):
    """The CUDA-graph replay of the T-step filter == infer_fused step by step (injected noise/uniforms)."""
    T, B, K = 9, 4, 512
    model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device=cuda)
    f = fused.GraphedFilter(model, T, B, K, inject_noise=True)
    gen = torch.Generator(device=cuda).manual_seed(1)
    for rep in range(2):
        obs = torch.randn(T, B, device=cuda, generator=gen)
        noise = torch.randn(T, B, K, device=cuda, generator=gen)
        u = torch.rand(T - 1, B, dtype=torch.float64, device=cuda, generator=gen)
        f.noise.copy_(noise)
        f.uniforms.copy_(u)
        got = f(obs.cpu())                      # host observations are copied in
        with torch.no_grad():
            ref = fused.infer_fused(model, obs, K, return_log_marginal_likelihood=True, return_latents=False,
                                    uniforms=u, noise=noise)
        assert torch.equal(got, ref["log_marginal_likelihood"])
        assert torch.equal(f.log_weight, ref["log_weight"]) and torch.equal(f.last_latent, ref["last_latent"])
    f.check()


def test_graphed_filter_fresh_randomness_and_kalman(cuda):
    T, B, K = 50, 8, 4096
    ys = lgssm.simulate(T, B, seed=4)
    exact = kalman.lgssm1d_log_evidence(ys, 0.0, 1.0, 0.9, 1.0, 1.0, 0.25)
    model = fused.ScalarLinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 0.0, 0.5, device=cuda)
    torch.manual_seed(0)
    f = fused.GraphedFilter(model, T, B, K)
    obs = torch.from_numpy(ys)
    runs = torch.stack([f(obs) for _ in range(8)]).cpu().numpy()
    assert len({r.tobytes() for r in runs}) == 8                # every replay uses new noise and uniforms
    assert np.abs(runs.mean(axis=0) - exact).max() < 0.5        # averaged over replays the bias is small
    f.check()
    # BASELINE config 1 shape (B = 1, K = 100, T = 50): one replay, no per-step host work
    small = fused.GraphedFilter(model, 50, 1, 100)
    out = small(torch.from_numpy(ys[:, :1]))
    assert out.shape == (1,) and torch.isfinite(out).all()


# ---- training through the fused kernels (aesmc_lg_step_bwd_f32) -------------------------------------------------
def _port_loss_and_grads(obs_np, K, u, seed):
    """losses.get_loss('aesmc') of the CPU oracle port (the reference's algorithm, torch autograd) on the
    reference-style trainable LGSSM, Normal.rsample fed from torch.Generator(seed); returns (loss, grads, noise)."""
    import torch.distributions.normal as tdn
    from oracle import reference_port as port
    torch.manual_seed(0)
    mods = (lgssm.Initial(0.1, 1.2), lgssm.Transition(0.8, 0.9), lgssm.Emission(1.1, 0.6), lgssm.Proposal(0.8, 0.7))
    gen = torch.Generator().manual_seed(seed)
    drawn = []
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        z = torch.randn(tuple(shape), generator=gen, dtype=dtype)
        drawn.append(z)
        return z.to(device)

    tdn._standard_normal = fake
    try:
        loss = port.get_loss([torch.from_numpy(o) for o in obs_np], K, "aesmc", *mods, uniforms=u)
    finally:
        tdn._standard_normal = orig
    loss.backward()
    grads = [p.grad.clone() for m in mods[1:] for p in m.parameters()]
    B = obs_np.shape[1]
    noise = torch.stack([z if tuple(z.shape) == (B, K) else z.t().contiguous() for z in drawn])  # t = 0 is drawn [K, B]
    return loss.item(), grads, noise, mods


@pytest.mark.parametrize("K,T,B", [(256, 6, 4), (1024, 5, 3)])
def test_fused_training_gradients_match_reference_port(cuda, K, T, B):
    """get_loss('aesmc') through the fused forward + backward kernels against the oracle port (CPU, torch autograd
    over the reference's algorithm) on identical noise and uniforms: loss within 1e-5, every parameter gradient of
    the reference's trainable LGSSM (test/models/lgssm.py:19-72) within 1e-4 relative."""
    obs_np = lgssm.simulate(T, B, seed=4)
    u = np.random.default_rng(5).random((T - 1, B))
    ref_loss, ref_grads, noise, _ = _port_loss_and_grads(obs_np, K, u, seed=6)
    torch.manual_seed(0)
    init, trans, emis, prop = (lgssm.Initial(0.1, 1.2), lgssm.Transition(0.8, 0.9).to(cuda), lgssm.Emission(1.1, 0.6).to(cuda),
                               lgssm.Proposal(0.8, 0.7).to(cuda))
    view = fused.link(init, trans, emis, prop)
    obs = torch.from_numpy(obs_np).to(cuda)
    assert fused.model_of(init, trans, emis, prop) is view and fused.applicable(view, obs, K, evidence_only=True)
    assert not fused.applicable(view, obs, K, evidence_only=False)   # other outputs: differentiable generic path
    lml = fused.evidence_with_grad(view, obs, K, uniforms=u, noise=noise.to(cuda))
    loss = -lml.mean()
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref_loss, rtol=1e-5)
    got = [p.grad.cpu() for m in (trans, emis, prop) for p in m.parameters()]
    for g, r in zip(got, ref_grads):
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=1e-4, atol=2e-6)


def test_get_loss_dispatches_linked_modules_to_fused_kernels(cuda):
    """losses.get_loss with linked reference-style modules: 2 T launches (T forward steps, T backward steps), gradients
    on the modules' own Parameters, and a few Adam steps reduce the loss on data from the true model."""
    from aesmc_b200 import losses
    T, B, K = 10, 32, 512
    obs = torch.from_numpy(lgssm.simulate(T, B, A=0.9, Q=1.0, C=1.0, R=0.25, seed=3)).to(cuda)
    torch.manual_seed(1)
    np.random.seed(1)
    init, trans, emis, prop = (lgssm.Initial(0.0, 1.0), lgssm.Transition(0.2, 1.0).to(cuda), lgssm.Emission(0.3, 0.5).to(cuda),
                               lgssm.Proposal(0.9, 0.9).to(cuda))
    fused.link(init, trans, emis, prop)
    params = [p for m in (trans, emis, prop) for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=0.05)
    first = None
    for it in range(40):
        opt.zero_grad()
        n0 = aesmc_b200._lib.launch_count()
        loss = losses.get_loss([obs[t] for t in range(T)], K, "aesmc", init, trans, emis, prop)
        assert aesmc_b200._lib.launch_count() - n0 == T
        loss.backward()
        assert aesmc_b200._lib.launch_count() - n0 == 2 * T
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
        opt.step()
        first = loss.item() if first is None else first
    assert loss.item() < first - 0.5, (first, loss.item())
    assert abs(trans.mult.item()) > 0.4 and abs(emis.mult.item()) > 0.5      # moving towards (0.9, 1.0)


@pytest.mark.parametrize("proposal", ["bootstrap", AFFINE])
def test_fused_backward_equals_generic_autograd(cuda, proposal):
    """Same model object, same injected noise and uniforms: the fused forward + backward against this repo's generic
    path (torch-eager callables + step kernel + torch autograd), including the bootstrap proposal whose
    reparameterised path lands on the prior's parameters."""
    import torch.distributions.normal as tdn
    T, B, K = 6, 4, 256
    model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device=cuda)
    names = ["m0", "a", "b", "c", "d"]
    leaves = [getattr(model, n).requires_grad_() for n in names]
    if model.prop is not None:
        for k in ("p0_y", "p0_off", "pt_x", "pt_y", "pt_off"):
            leaves.append(model.prop[k].requires_grad_())
    obs = torch.from_numpy(lgssm.simulate(T, B, seed=1)).to(cuda)
    noise = torch.randn(T, B, K, device=cuda, generator=torch.Generator(device=cuda).manual_seed(0))
    u = np.random.default_rng(0).random((T - 1, B))
    lml = fused.evidence_with_grad(model, obs, K, uniforms=u, noise=noise)
    (-lml.mean()).backward()
    got = [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    it = iter(noise)
    orig = tdn._standard_normal
    tdn._standard_normal = lambda shape, dtype, device: (lambda z: z if tuple(shape) == tuple(z.shape) else z.t().contiguous())(next(it))
    try:
        res = inference.infer("smc", obs, *model.callables(), K, return_log_marginal_likelihood=True, return_latents=False,
                              return_log_weight=False, uniforms=u, _allow_fused=False)
    finally:
        tdn._standard_normal = orig
    assert torch.equal(res["log_marginal_likelihood"], lml.detach())
    (-res["log_marginal_likelihood"].mean()).backward()
    for g, t, n in zip(got, leaves, names + ["p0_y", "p0_off", "pt_x", "pt_y", "pt_off"]):
        np.testing.assert_allclose(g.cpu().numpy(), t.grad.cpu().numpy(), rtol=2e-4, atol=2e-6, err_msg=n)


@pytest.mark.parametrize("proposal", ["bootstrap", AFFINE])
def test_fused_scalar_matches_cpu_port_on_shared_noise(cuda, proposal):
    """Direct parity of the fused scalar path with the CPU port of the reference (inference.py:85-134 with torch CPU
    distributions and numpy resampling): same injected normals and uniforms on both sides."""
    import torch.distributions.normal as tdn
    from oracle import reference_port as port
    T, B, K = 8, 6, 1024
    obs_np = lgssm.simulate(T, B, seed=4)
    noise = torch.randn(T, B, K, generator=torch.Generator().manual_seed(1))
    u = np.random.default_rng(2).random((T - 1, B))
    kw = dict(return_log_marginal_likelihood=True, return_log_weights=True, return_ancestral_indices=True,
              return_original_latents=True)
    cpu_model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device="cpu")
    it = iter(noise)
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        z = next(it)
        return z if tuple(shape) == tuple(z.shape) else z.t().contiguous()

    tdn._standard_normal = fake
    try:
        with torch.no_grad():
            ref = port.infer("smc", [torch.from_numpy(o) for o in obs_np], *cpu_model.callables(), K, uniforms=u, **kw)
    finally:
        tdn._standard_normal = orig
    model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device=cuda)
    with torch.no_grad():
        got = fused.infer_fused(model, torch.from_numpy(obs_np).to(cuda), K, uniforms=u, noise=noise.to(cuda), **kw)
    anc_ref = torch.stack(ref["ancestral_indices"]).numpy()
    anc_got = torch.stack(got["ancestral_indices"]).cpu().numpy()
    lw_ref = torch.stack(ref["log_weights"]).numpy()
    lw_got = torch.stack(got["log_weights"]).cpu().numpy()
    same_rows = (anc_ref == anc_got).all(axis=(0, 2))
    print("fused scalar vs CPU port (%s): ancestors differing %d of %d, rows with identical genealogy %d of %d, "
          "log-weight bits differing at t = 0: %d" % ("bootstrap" if proposal == "bootstrap" else "affine",
                                                     (anc_ref != anc_got).sum(), anc_ref.size, same_rows.sum(), B,
                                                     (lw_ref[0].view(np.int32) != lw_got[0].view(np.int32)).sum()))
    rel = np.abs(lw_got[0] - lw_ref[0]) / np.maximum(np.abs(lw_ref[0]), 1.0)
    assert rel.max() < 1e-5                                   # north-star tolerance for log-weights
    assert (anc_ref[0] != anc_got[0]).mean() < 2e-3           # first resampling step: same inputs on both sides
    assert same_rows.sum() >= B - 2
    dz = np.abs(got["log_marginal_likelihood"].cpu().numpy() - ref["log_marginal_likelihood"].numpy())
    assert dz[same_rows].max() < 1e-3 and dz.max() < 1.0


@pytest.mark.parametrize("proposal", ["bootstrap", AFFINE])
def test_fused_step_rare_paths_forced(cuda, proposal):
    """The fused-model instance of the exact row kernel with its two rare paths forced on every row (test hook
    aesmc_debug_force_rare_paths: sequential redo after a failed verification, float64 redo of the run marks): the
    filter's outputs keep their bits."""
    from aesmc_b200 import _lib
    T, B, K = 6, 9, 4096
    model = fused.ScalarLinearGaussianSSM(0.2, 1.1, 0.9, 0.05, 0.7, 1.3, -0.1, 0.5, proposal=proposal, device=cuda)
    obs = torch.from_numpy(lgssm.simulate(T, B, seed=3)).to(cuda)
    noise = torch.randn(T, B, K, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    u = np.random.default_rng(1).random((T - 1, B))
    kw = dict(return_log_marginal_likelihood=True, return_latents=True, return_log_weights=True, return_ancestral_indices=True)
    with torch.no_grad():
        ref = fused.infer_fused(model, obs, K, uniforms=u, noise=noise, **kw)
    lib = _lib.load()
    for forced in (1, 2, 3):
        prev = lib.aesmc_debug_force_rare_paths(forced)
        try:
            with torch.no_grad():
                got = fused.infer_fused(model, obs, K, uniforms=u, noise=noise, **kw)
            torch.cuda.synchronize()
        finally:
            lib.aesmc_debug_force_rare_paths(prev)
        assert torch.equal(got["log_marginal_likelihood"], ref["log_marginal_likelihood"]), forced
        for a, b in zip(got["ancestral_indices"], ref["ancestral_indices"]):
            assert torch.equal(a, b), forced
        for t in range(T):
            assert torch.equal(got["latents"][t], ref["latents"][t]) and torch.equal(got["log_weights"][t], ref["log_weights"][t])
