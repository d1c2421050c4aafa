"""Fused path for vector latents (BASELINE config 3; aesmc_lgv_propose_f32 + the step kernel): D-dimensional
linear-Gaussian models with diagonal noise.  Checked against the repo's generic path (torch-eager callables), against
the CPU port of the reference on the same injected normals and uniforms, against the C oracle's resampling on the
kernel's own log-weights (indices bit-exact), and against the matrix Kalman filter."""
import numpy as np
import pytest
import torch

import aesmc_b200
from aesmc_b200 import fused, inference
from oracle import core as oracle
from oracle import kalman
from tests.models import lgssm_dense

pytestmark = pytest.mark.gpu

KW = dict(return_log_marginal_likelihood=True, return_latents=True, return_original_latents=True, return_log_weight=True,
          return_log_weights=True, return_ancestral_indices=True)


def make_model(dev, D, Dy, proposal, seed=0):
    g = torch.Generator().manual_seed(seed)
    A, C = lgssm_dense.make_system(D, Dy, seed=seed + 1)
    rnd = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    prop = "bootstrap"
    if proposal == "learned":
        prop = {"W0": 0.3 * rnd(D, Dy), "b0": 0.1 * rnd(D), "s0": 0.8 + 0.2 * torch.rand(D, generator=g),
                "Wx": 0.5 * A + 0.05 * rnd(D, D), "Wy": 0.2 * rnd(D, Dy), "bt": 0.1 * rnd(D), "st": 0.5 + 0.2 * torch.rand(D, generator=g)}
    return fused.VectorLinearGaussianSSM(0.1 * rnd(D), 1.0 + 0.1 * torch.rand(D, generator=g), A, 0.5, C, 0.7, b=0.05 * rnd(D),
                                         d=0.05 * rnd(Dy), proposal=prop, device=dev)


def with_noise(fn, model, obs, K, noise, u, hide=True, **kw):
    """Run `fn` (an infer) on the model's callables with every Normal.rsample fed from noise[t] ([B, K, D]; transposed
    for batch-expanded distributions, which are sampled as [K, B, D])."""
    import torch.distributions.normal as tdn
    it = iter(noise)
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        z = next(it).to(device)
        return z if tuple(shape) == tuple(z.shape) else z.transpose(0, 1).contiguous()

    tdn._standard_normal = fake
    try:
        with torch.no_grad():
            cs = model.callables()
            if hide:  # the lambdas hide the bound methods, so infer() takes the generic path
                cs = (lambda: model.initial(), lambda **k: model.transition(**k), lambda **k: model.emission(**k),
                      lambda **k: model.proposal(**k))
            return fn("smc", obs, *cs, K, uniforms=u, **kw)
    finally:
        tdn._standard_normal = orig


@pytest.mark.parametrize("D,Dy", [(10, 10), (3, 2), (16, 5), (1, 1)])
@pytest.mark.parametrize("proposal", ["bootstrap", "learned"])
def test_vector_fused_matches_generic_path(cuda, D, Dy, proposal):
    T, B, K = 6, 5, 1024
    model = make_model(cuda, D, Dy, proposal)
    gen = torch.Generator(device=cuda).manual_seed(0)
    obs = torch.randn(T, B, Dy, device=cuda, generator=gen)
    noise = torch.randn(T, B, K, D, device=cuda, generator=gen)
    u = np.random.default_rng(0).random((T - 1, B))
    ref = with_noise(inference.infer, model, obs, K, noise, u, **KW)
    with torch.no_grad():
        got = fused.infer_fused_vector(model, obs, K, uniforms=u, noise=noise, **KW)
    # before any resampling the two paths see the same inputs: proposed latents and log-weights agree to rounding
    assert torch.allclose(got["original_latents"][0], ref["original_latents"][0], rtol=1e-5, atol=1e-6)
    lw0, lr0 = got["log_weights"][0], ref["log_weights"][0]
    assert ((lw0 - lr0).abs() / lr0.abs().clamp_min(1.0)).max().item() < 1e-5   # 1e-5 relative (north star)
    # the resampling itself is exact: the C oracle on the kernel's OWN log-weights gives the kernel's ancestors
    for t in range(T - 1):
        lw = got["log_weights"][t].cpu().numpy()
        want, st = oracle.sample_ancestral_index(lw, u[t])[:2]
        assert st == 0 and np.array_equal(got["ancestral_indices"][t].cpu().numpy(), want), t
        assert torch.equal(got["latents"][-1], got["original_latents"][-1])
    # Step by step on the GENERIC path's inputs (teacher forcing: once a single ancestor differs the two filters hold
    # different particles and are only statistically comparable): the model kernel on the reference's resampled
    # latents reproduces the reference's proposals and log-weights to rounding at every time step.
    p0, pt = model.kernel_params()
    q_rows = model.proposal_row_means(obs)
    for t in range(T):
        x_prev = None
        if t:
            a = ref["ancestral_indices"][t - 1]
            x_prev = torch.gather(ref["original_latents"][t - 1], 1, a.unsqueeze(-1).expand(-1, -1, D)).contiguous()
        x_new, lw = torch.empty(B, K, D, device=cuda), torch.empty(B, K, device=cuda)
        aesmc_b200._lib.call("aesmc_lgv_propose_f32", aesmc_b200._lib.ptr(x_prev), aesmc_b200._lib.ptr(obs[t].contiguous()),
                             aesmc_b200._lib.ptr(noise[t].contiguous()),
                             aesmc_b200._lib.ptr(None if q_rows is None else q_rows[t]), (p0 if t == 0 else pt).ctypes.data,
                             D, Dy, int(q_rows is None), 0, t, B, K, aesmc_b200._lib.ptr(x_new), aesmc_b200._lib.ptr(lw))
        assert torch.allclose(x_new, ref["original_latents"][t], rtol=1e-5, atol=2e-6), t
        lr = ref["log_weights"][t]
        assert ((lw - lr).abs() / lr.abs().clamp_min(1.0)).max().item() < 1e-5, t
    # end to end: ancestors of the first resampling step (same inputs on both sides) differ only where a log-weight
    # differs in its last bits; the evidence agrees row for row where no ancestor differs, statistically elsewhere
    first = (got["ancestral_indices"][0] != ref["ancestral_indices"][0])
    print("D=%d Dy=%d %s: ancestors differing at t = 0: %d of %d" % (D, Dy, proposal, int(first.sum()), first.numel()))
    assert first.float().mean().item() < 5e-3
    same_rows = torch.stack([(a == b).all(dim=1) for a, b in zip(got["ancestral_indices"], ref["ancestral_indices"])]).all(dim=0)
    dz = (got["log_marginal_likelihood"] - ref["log_marginal_likelihood"]).abs()
    if same_rows.any():
        assert dz[same_rows].max().item() < 2e-3
    assert dz.max().item() < 1.0
    assert got["last_latent"].shape == (B, K, D) and got["log_weight"].shape == (B, K)


def test_vector_fused_matches_cpu_port_on_shared_noise(cuda):
    """The same normals and uniforms through the CPU port of the reference (torch CPU model, numpy resampling) and the
    fused GPU path."""
    from oracle import reference_port as port
    T, B, K, D, Dy = 6, 4, 1024, 10, 10
    gen = torch.Generator().manual_seed(3)
    obs = torch.randn(T, B, Dy, generator=gen)
    noise = torch.randn(T, B, K, D, generator=gen)
    u = np.random.default_rng(1).random((T - 1, B))
    ref = with_noise(port.infer, make_model("cpu", D, Dy, "learned"), list(obs), K, noise, u, hide=False, **KW)
    with torch.no_grad():
        got = fused.infer_fused_vector(make_model(cuda, D, Dy, "learned"), obs.to(cuda), K, uniforms=u, noise=noise.to(cuda), **KW)
    lw0, lr0 = got["log_weights"][0].cpu(), ref["log_weights"][0]
    assert ((lw0 - lr0).abs() / lr0.abs().clamp_min(1.0)).max().item() < 1e-5
    anc_ref = torch.stack(ref["ancestral_indices"]).numpy()
    anc_got = torch.stack(got["ancestral_indices"]).cpu().numpy()
    first = anc_ref[0] != anc_got[0]
    print("10-D fused vs CPU port: ancestors differing at t = 0: %d of %d" % (first.sum(), first.size))
    assert first.mean() < 5e-3     # same inputs on both sides; later steps are comparable only where nothing differed before
    same_rows = (anc_ref == anc_got).all(axis=(0, 2))
    dz = np.abs(got["log_marginal_likelihood"].cpu().numpy() - ref["log_marginal_likelihood"].numpy())
    print("rows with identical genealogy: %d of %d, max |d log Z| there %.2e, overall %.2e"
          % (same_rows.sum(), B, dz[same_rows].max() if same_rows.any() else 0.0, dz.max()))
    if same_rows.any():
        assert dz[same_rows].max() < 2e-3
    assert dz.max() < 1.0


def test_infer_dispatches_linked_dense_modules_and_tracks_kalman(cuda):
    T, B, K, dx, dy = 10, 6, 16384, 10, 10
    s0, q, r = 1.0, 0.5, 2.0
    A, C = lgssm_dense.make_system(dx, dy, seed=1, device=cuda)
    init, trans, emis = lgssm_dense.Initial(dx, s0, cuda), lgssm_dense.Transition(A, q), lgssm_dense.Emission(C, r)
    prop = lgssm_dense.PriorProposal(init, trans)
    ys = lgssm_dense.simulate(A, C, T, B, s0, q, r, seed=2)
    exact = kalman.lgssm_log_evidence(ys.numpy(), np.zeros(dx), s0 ** 2 * np.eye(dx), A.cpu().numpy(), q ** 2 * np.eye(dx),
                                      C.cpu().numpy(), r ** 2 * np.eye(dy))
    view = fused.link_dense(init, trans, emis, prop)
    assert fused.model_of(init, trans, emis, prop) is view and view.prop is None
    obs = [y.to(cuda) for y in ys]
    launches = aesmc_b200._lib.launch_count()
    torch.manual_seed(0)
    np.random.seed(0)
    with torch.no_grad():
        res = inference.infer("smc", obs, init, trans, emis, prop, K, return_log_marginal_likelihood=True, return_latents=False)
    assert aesmc_b200._lib.launch_count() - launches == 2 * T     # one model launch + one step launch per time step
    err = np.abs(res["log_marginal_likelihood"].cpu().numpy() - exact)
    print("10-D fused bootstrap filter: max |log Z_hat - log Z| =", err.max(), "at K =", K)
    assert err.max() < 1.0
    # the learned proposal of config 3 links too, and a model that needs gradients takes the generic path
    torch.manual_seed(0)
    learned = lgssm_dense.Proposal(dx, dy).to(cuda)
    v2 = fused.link_dense(init, trans, emis, learned)
    assert fused.model_of(init, trans, emis, learned) is v2 and set(v2.prop) == {"W0", "b0", "s0", "Wx", "Wy", "bt", "st"}
    assert not fused.applicable(v2, obs, 1024)          # parameters require grad and grad mode is on
    with torch.no_grad():
        assert fused.applicable(v2, obs, 1024)
        r1 = inference.infer("smc", obs, init, trans, emis, learned, 1024, return_log_marginal_likelihood=True)
    assert r1["latents"][0].shape == (B, 1024, dx) and torch.isfinite(r1["log_marginal_likelihood"]).all()
    with pytest.raises(ValueError):
        fused.link_dense(init, trans, emis, object())


def test_vector_philox_noise_is_standard_and_seeded(cuda):
    model = make_model(cuda, 10, 10, "bootstrap")
    obs = torch.zeros(1, 8, 10, device=cuda)
    torch.manual_seed(3)
    with torch.no_grad():
        x = fused.infer_fused_vector(model, obs, 16384, return_latents=False)["last_latent"]
        z = ((x - model.m0) / model.s0).double().flatten()
    n = z.numel()
    assert abs(z.mean().item()) < 5 / np.sqrt(n) and abs(z.var().item() - 1) < 5 * np.sqrt(2 / n)
    assert abs((z ** 4).mean().item() - 3) < 0.05 and z.unique().numel() > 0.99 * n
    torch.manual_seed(3)
    with torch.no_grad():
        x2 = fused.infer_fused_vector(model, obs, 16384, return_latents=False)["last_latent"]
    assert torch.equal(x, x2)
