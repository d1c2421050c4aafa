"""GPU tests of the remaining C-ABI entry points against the oracle / plain torch: gather and its
backward, logsumexp, step backward, log_ess, weighted moments, index utilities, and the public
math / state / statistics wrappers (including the reference's own known-answer vectors)."""
import numpy as np
import pytest
import torch

import aesmc_b200
from aesmc_b200 import _lib, _ops, inference, math as amath, state, statistics
from oracle import core as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,dtype", [((3, 7), torch.float32), ((4, 100, 3), torch.float32),
                                         ((2, 513, 10), torch.float32), ((2, 64, 4, 5), torch.float64),
                                         ((3, 33, 3), torch.int64), ((2, 50, 3), torch.uint8),
                                         ((2, 40, 5), torch.float16)])
def test_gather_matches_torch(cuda, shape, dtype):
    g = torch.Generator().manual_seed(1)
    B, K = shape[:2]
    v = (torch.randn(shape, generator=g) * 10).to(dtype).to(cuda)
    idx = torch.randint(0, K, (B, K), generator=g).to(cuda)
    want = torch.gather(v, 1, idx.reshape(B, K, *([1] * (v.dim() - 2))).expand_as(v))
    assert torch.equal(state.resample(v, idx), want)
    assert torch.equal(state.resample(v, idx.int()), want)
    assert torch.equal(state.resample({"a": v}, idx)["a"], want)


def test_resample_known_answers_and_cpu_staging(cuda):
    # test/test_state.py:286-303 on CPU tensors (staged through the GPU, returned on the CPU)
    value = torch.Tensor([[1, 2, 3], [4, 5, 6]])
    idx = torch.LongTensor([[1, 2, 0], [0, 0, 1]])
    out = state.resample(value, idx)
    assert not out.is_cuda and torch.equal(out, torch.Tensor([[2, 3, 1], [4, 4, 5]]))
    assert state.resample(torch.rand(3, 2, 4, 5), torch.zeros(3, 2).long()).shape == (3, 2, 4, 5)
    with pytest.raises(AttributeError):
        state.resample([1, 2], idx)
    with pytest.raises(IndexError):
        state.resample(value, torch.LongTensor([[1, 2, 3], [0, 0, 1]]))


def test_genealogy_known_answer(cuda):
    # test/test_inference.py:13-40
    latents = [torch.Tensor([[1, 2, 3]]), torch.Tensor([[4, 5, 6]]), torch.Tensor([[7, 8, 9]]), torch.Tensor([[10, 11, 12]])]
    anc = [torch.LongTensor([[0, 2, 1]]), torch.LongTensor([[2, 0, 0]]), torch.LongTensor([[1, 2, 0]])]
    want = [[1, 1, 2], [4, 4, 6], [8, 9, 7], [10, 11, 12]]
    for dev in ("cpu", cuda):
        got = inference.get_resampled_latents([l.to(dev) for l in latents], [a.to(dev) for a in anc])
        assert [g[0].tolist() for g in got] == want
    rng = np.random.default_rng(0)
    T, B, K = 6, 3, 50
    lat = [torch.from_numpy(rng.standard_normal((B, K, 2)).astype(np.float32)).to(cuda) for _ in range(T)]
    anc = [torch.from_numpy(np.sort(rng.integers(0, K, (B, K)), axis=1)).to(cuda) for _ in range(T - 1)]
    got = inference.get_resampled_latents(lat, anc)
    cur = np.tile(np.arange(K), (B, 1))
    for t in range(T - 1, -1, -1):
        assert np.array_equal(got[t].cpu().numpy(), oracle.resample(lat[t].cpu().numpy(), cur))
        if t:
            cur = oracle.compose_index(anc[t - 1].cpu().numpy(), cur)


@pytest.mark.parametrize("sorted_rows", [True, False])
@pytest.mark.parametrize("D", [1, 2, 3, 4, 8, 10])
def test_gather_backward(cuda, sorted_rows, D):
    rng = np.random.default_rng(3)
    B, K = 6, 701
    idx = rng.integers(0, K, (B, K))
    if sorted_rows:
        idx = np.sort(idx, axis=1)
        idx[0] = 17                               # one parent takes everything
        idx[1] = np.arange(K)                     # every parent exactly one child
        idx[2, : K // 2] = 0                      # long run at the start, childless parents after it
        idx[3] = np.sort(rng.integers(K - 3, K, K))  # everything at the end of the row
    x = torch.from_numpy(rng.standard_normal((B, K, D)).astype(np.float32)).to(cuda).requires_grad_()
    g = torch.from_numpy(rng.standard_normal((B, K, D)).astype(np.float32)).to(cuda)
    out = _ops.gather(x, torch.from_numpy(idx).int().to(cuda), sorted_rows=sorted_rows)
    out.backward(g)
    ref = oracle.resample_bwd(g.cpu().numpy(), idx)
    got = x.grad.cpu().numpy()
    if sorted_rows:
        # runs of up to 16 children are summed in k order == the reference's CPU scatter_add order, bit for bit;
        # longer runs (collapsed weights) are finished by a warp-wide fixed-shape reduction
        short = np.array([np.bincount(idx[b], minlength=K).max() <= 16 for b in range(B)])
        assert short.sum() >= 3 and (~short).sum() >= 3
        assert np.array_equal(got[short], ref[short])
        np.testing.assert_allclose(got[~short], ref[~short], rtol=1e-5, atol=2e-5)
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("K", [128, 1024, 4096, 16384])
@pytest.mark.parametrize("wide", [False, True])
def test_gather_backward_row_kernel(cuda, K, wide):
    """Scalar latents, K % 4 == 0, K <= 16384: the parent-centric row kernel (gather_bwd_rows.cu).  Runs of up to 32
    children are summed in particle order == the reference's CPU scatter_add, bit for bit; longer ones by a warp."""
    rng = np.random.default_rng(K)
    B = 9
    idx = np.sort(rng.integers(0, K, (B, K)), axis=1)
    idx[0] = 17                                   # one parent takes everything
    idx[1] = np.arange(K)                         # every parent exactly one child
    idx[2, : K // 2] = 0                          # long run at the start, childless parents after it
    idx[3] = np.sort(rng.integers(K - 3, K, K))   # everything at the end of the row
    idx[4] = np.repeat(np.arange(K // 32), 32)    # runs of exactly 32: the longest a thread sums itself
    idx[5] = np.sort(np.concatenate([np.repeat(np.arange(0, K, K // 4)[:3], 33), rng.integers(0, K, K - 99)]))  # runs just over
    g = torch.from_numpy(rng.standard_normal((B, K)).astype(np.float32)).to(cuda)
    x = torch.zeros(B, K, device=cuda, requires_grad=True)
    it = torch.from_numpy(idx).to(cuda)
    _ops.gather(x, it if wide else it.int(), sorted_rows=True).backward(g)
    ref = oracle.resample_bwd(g.cpu().numpy(), idx)
    got = x.grad.cpu().numpy()
    short = np.array([np.bincount(idx[b], minlength=K).max() <= 32 for b in range(B)])
    assert short.sum() >= 4 and (~short).sum() >= 3
    assert np.array_equal(got[short], ref[short])
    np.testing.assert_allclose(got[~short], ref[~short], rtol=1e-5, atol=1e-4 * np.sqrt(K / 1024))
    # in the long-run rows every parent with a short run is still exact
    for b in np.nonzero(~short)[0]:
        counts = np.bincount(idx[b], minlength=K)
        assert np.array_equal(got[b][counts <= 32], ref[b][counts <= 32])


@pytest.mark.parametrize("D", [1, 2, 5])
def test_gather_backward_long_rows(cuda, D):
    """Sorted backward on long rows: runs that span many threads and CTAs, odd K (no vector alignment)."""
    rng = np.random.default_rng(D)
    B, K = 2, 30011
    idx = np.sort(rng.integers(0, K, (B, K)), axis=1)
    idx[1, 100:9000] = idx[1, 100]               # a run spanning many threads
    g = torch.from_numpy(rng.standard_normal((B, K, D)).astype(np.float32)).to(cuda)
    x = torch.zeros(B, K, D, device=cuda, requires_grad=True)
    _ops.gather(x, torch.from_numpy(idx).int().to(cuda), sorted_rows=True).backward(g)
    ref = oracle.resample_bwd(g.cpu().numpy(), idx)
    assert np.array_equal(x.grad.cpu().numpy()[0], ref[0])          # short runs: the reference's order, bit for bit
    np.testing.assert_allclose(x.grad.cpu().numpy()[1], ref[1], rtol=1e-5, atol=1e-4)  # the 8 900-child run: warp reduction


def test_step_backward_matches_torch_autograd(cuda):
    gen = torch.Generator(device=cuda).manual_seed(2)
    B, K, D = 6, 1000, 3
    leaves = [torch.randn(B, K, device=cuda, generator=gen).requires_grad_() for _ in range(3)]
    x = torch.randn(B, K, D, device=cuda, generator=gen).requires_grad_()
    u = torch.rand(B, dtype=torch.float64, device=cuda, generator=gen)
    flags = _ops.new_flags(cuda)
    log_w, lse, idx, xr = _ops.smc_step(*leaves, u, x, flags, "exact", True)
    cot = [torch.randn_like(log_w), torch.randn_like(lse), torch.randn_like(xr)]
    (log_w * cot[0]).sum().add((lse * cot[1]).sum()).add((xr * cot[2]).sum()).backward()
    got = [t.grad.clone() for t in leaves + [x]]
    for t in leaves + [x]:
        t.grad = None
    lw2 = (leaves[0] + leaves[1]) - leaves[2]
    lse2 = torch.logsumexp(lw2, dim=1)
    xr2 = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(B, K, D))
    (lw2 * cot[0]).sum().add((lse2 * cot[1]).sum()).add((xr2 * cot[2]).sum()).backward()
    for mine, t in zip(got, leaves + [x]):
        torch.testing.assert_close(mine, t.grad, rtol=1e-5, atol=1e-6)


def test_logsumexp_and_log_ess(cuda, golden):
    g = golden["default"]
    lw = torch.from_numpy(g["stats/lw"]).to(cuda)
    np.testing.assert_allclose(statistics.log_ess(lw).cpu().numpy(), g["stats/log_ess"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(statistics.ess(lw).cpu().numpy(), g["stats/ess"], rtol=1e-5)
    rng = np.random.default_rng(0)
    for B, K in ((1, 1), (3, 5), (40, 1000), (7, 5000)):
        a = (rng.standard_normal((B, K)) * 4).astype(np.float32)
        np.testing.assert_allclose(_ops.logsumexp_rows(torch.from_numpy(a).to(cuda)).cpu().numpy(), oracle.lse_f64(a), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(statistics.log_ess(torch.from_numpy(a).to(cuda)).cpu().numpy(), oracle.log_ess_f64(a), rtol=1e-5, atol=1e-5)
    # shapes: test/test_statistics.py:46-69
    assert statistics.log_ess(-torch.rand(2, 3)).shape == (2,)
    assert statistics.log_ess(-torch.rand(3, 1)).shape == (3,)
    assert statistics.log_ess(-torch.rand(3)).shape == ()
    # float64 with +-1e6 offsets: test/test_statistics.py:71-115
    nw = np.array([0.2, 0.3, 0.5])
    for shift in (np.log(0.47), 1e6, -1e6):
        lw64 = torch.from_numpy(np.log(nw) + shift)
        assert abs(statistics.log_ess(lw64).item() - np.log(1 / np.sum(nw ** 2))) < 1e-7
        assert abs(statistics.ess(lw64.to(cuda)).item() - 1 / np.sum(nw ** 2)) < 1e-7
    # differentiable path
    t = torch.randn(4, 50, device=cuda, requires_grad=True)
    statistics.log_ess(t).sum().backward()
    t2 = t.detach().clone().requires_grad_()
    (2 * torch.logsumexp(t2, 1) - torch.logsumexp(2 * t2, 1)).sum().backward()
    torch.testing.assert_close(t.grad, t2.grad, rtol=1e-5, atol=1e-6)


def test_math_surface(cuda):
    # shapes / dims / type preservation: test/test_math.py:9-49
    for values in (torch.rand(2, 3, 4, 5), np.random.rand(2, 3, 4, 5)):
        for fn in (amath.lognormexp, amath.exponentiate_and_normalize):
            out = fn(values, dim=2)
            assert type(out) is type(values) and tuple(out.shape) == (2, 3, 4, 5)
    x = torch.Tensor([1, 2, 3])
    want = torch.exp(x) / torch.exp(x).sum()
    np.testing.assert_allclose(amath.lognormexp(x).numpy(), torch.log(want).numpy(), atol=1e-6)
    np.testing.assert_allclose(amath.exponentiate_and_normalize(x).numpy(), want.numpy(), rtol=1e-6)
    np.testing.assert_allclose(amath.exponentiate_and_normalize(np.array([1., 2., 3.])),
                               np.exp([1., 2., 3.]) / np.exp([1., 2., 3.]).sum(), rtol=1e-7)
    v = torch.randn(5, 6, 7, device=cuda)
    torch.testing.assert_close(amath.lognormexp(v, dim=1), v - torch.logsumexp(v, 1, keepdim=True), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(amath.exponentiate_and_normalize(v, dim=0), torch.softmax(v, 0), rtol=1e-5, atol=1e-7)
    r = torch.randn(3, 9, device=cuda, requires_grad=True)
    amath.lognormexp(r, dim=1)[:, 0].sum().backward()
    r2 = r.detach().clone().requires_grad_()
    torch.log_softmax(r2, 1)[:, 0].sum().backward()
    torch.testing.assert_close(r.grad, r2.grad, rtol=1e-5, atol=1e-6)


def test_statistics_moments(cuda, golden):
    g = golden["default"]
    lw, val = torch.from_numpy(g["stats/lw"]), torch.from_numpy(g["stats/value"])
    for dev in ("cpu", cuda):
        np.testing.assert_allclose(statistics.empirical_mean(val.to(dev), lw.to(dev)).cpu().numpy(), g["stats/mean"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(statistics.empirical_variance(val.to(dev), lw.to(dev)).cpu().numpy(), g["stats/var"], rtol=1e-4, atol=1e-5)
    # general f keeps the reference's particle loop: test/test_statistics.py:9-42
    out = statistics.empirical_expectation(torch.rand(2, 3, 4, 5, 6), -torch.rand(2, 3), lambda a: torch.rand(2, 7, 8))
    assert out.shape == (2, 7, 8)
    value = torch.Tensor([1, 2, 3]).unsqueeze(0)
    lw3 = torch.log(torch.Tensor([0.2, 0.3, 0.5]).unsqueeze(0))
    got = statistics.empirical_expectation(value, lw3, lambda v: v * 2)
    np.testing.assert_allclose(got.numpy(), [1 * 2 * 0.2 + 2 * 2 * 0.3 + 3 * 2 * 0.5], rtol=1e-6)
    # scalar latents [B,K]
    x = torch.randn(6, 300, device=cuda)
    w = torch.randn(6, 300, device=cuda)
    torch.testing.assert_close(statistics.empirical_mean(x, w), (torch.softmax(w, 1) * x).sum(1), rtol=1e-5, atol=1e-6)


def test_sample_ancestral_index_surface(cuda):
    # test/test_inference.py:44-84
    for shape in ((2, 3), (1, 2), (2, 1)):
        assert inference.sample_ancestral_index(torch.rand(*shape)).size() == torch.Size(shape)
    out = inference.sample_ancestral_index(torch.rand(1, 1))
    assert out.dtype == torch.int64 and not out.is_cuda
    assert inference.sample_ancestral_index(torch.rand(2, 5, device=cuda)).is_cuda
    weight = [0.2, 0.3, 0.5]
    idx = inference.sample_ancestral_index(torch.log(torch.Tensor(weight)).unsqueeze(0).expand(10000, 3))
    freq = np.bincount(idx.numpy().ravel(), minlength=3) / idx.numel()
    np.testing.assert_allclose(freq, weight, atol=1e-2)
    with pytest.raises(FloatingPointError):
        inference.sample_ancestral_index(torch.Tensor([[0.0, float("nan")]]))
    # same numpy seed -> same draws as the reference's np.random.uniform(size=[B,1])
    lw = torch.randn(7, 50)
    np.random.seed(5)
    a = inference.sample_ancestral_index(lw)
    np.random.seed(5)
    u = np.random.uniform(size=[7, 1])
    ref, _ = oracle.sample_ancestral_index(lw.numpy(), u)
    assert np.array_equal(a.numpy(), ref)


def test_index_utils(cuda):
    i = torch.randint(0, 1000, (5, 77), device=cuda)
    assert torch.equal(_ops.widen_index(_ops.narrow_index(i)), i)
    assert torch.equal(_ops.iota_index(3, 9, cuda).long(), torch.arange(9, device=cuda).expand(3, 9))


def test_specialised_expf_exhaustive(cuda):
    """np_expf_nonpos (hot-path exp: custom correctly-rounded division, no range tests) equals the
    general reference-order np_expf for EVERY float in [-104, -0] and -inf (1.12e9 inputs, on device);
    np_expf itself is pinned to numpy's bits through the oracle in test_step_parity_gpu.py."""
    from aesmc_b200 import _lib
    out = torch.zeros(2, dtype=torch.int64, device=cuda)
    _lib.call("aesmc_selftest_expf", out.data_ptr())
    torch.cuda.synchronize()
    mism, where = (int(v) for v in out.cpu())
    assert mism == 0, "np_expf_nonpos differs from np_expf on %d inputs, e.g. bits 0x%08x" % (mism, where)


def test_fused_normal_log_prob_bitwise_and_gradients(cuda):
    """state.log_prob's one-kernel Normal path == torch.distributions.Normal.log_prob bit for bit, for every
    operand layout the SMC driver produces; gradients match torch autograd."""
    from aesmc_b200 import _ops
    B, K = 7, 300
    gen = torch.Generator(device=cuda).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=cuda, generator=gen)  # noqa: E731
    Normal = torch.distributions.Normal
    FULL, BATCH = state.BatchShapeMode.FULLY_EXPANDED, state.BatchShapeMode.BATCH_EXPANDED
    cases = []
    v = rnd(B, K)
    cases.append(("dense loc, python scale", Normal(rnd(B, K), 0.7), v, FULL))
    cases.append(("dense loc, cuda 0-dim scale", Normal(rnd(B, K), torch.tensor(1.3, device=cuda)), v, FULL))
    cases.append(("cuda scalars", Normal(torch.tensor(0.2, device=cuda), torch.tensor(1.1, device=cuda)), v, None))
    cases.append(("cpu scalars", Normal(0.3, 2.0), v, None))
    cases.append(("per-row loc (batch expanded)", Normal(rnd(B), 0.9), v, BATCH))
    cases.append(("expanded observation", Normal(rnd(B, K), 0.5), state.expand_observation(rnd(B), K), FULL))
    for name, d, value, mode in cases:
        if mode is not None:
            state.set_batch_shape_mode(d, mode)
        got = state.log_prob(d, value)
        if mode is BATCH:
            want = d.log_prob(value.transpose(0, 1)).transpose(0, 1)
        else:
            want = d.log_prob(value)
        assert got.shape == (B, K) and torch.equal(got, want.expand(B, K)), name
        assert _ops.normal_log_prob(d, value) is not None, name
    # gradients w.r.t. value, dense loc, per-row loc and a CUDA scalar scale
    for loc_shape in ((B, K), (B,)):
        leaves = [rnd(B, K).requires_grad_(), rnd(*loc_shape).requires_grad_(), torch.tensor(0.8, device=cuda, requires_grad=True)]
        g = rnd(B, K)

        def run(fn):
            for t in leaves:
                t.grad = None
            d = Normal(leaves[1], leaves[2])
            if loc_shape == (B,):
                state.set_batch_shape_mode(d, BATCH)
            fn(d).mul(g).sum().backward()
            return [t.grad.clone() for t in leaves]

        mine = run(lambda d: state.log_prob(d, leaves[0]))
        if loc_shape == (B,):
            ref = run(lambda d: d.log_prob(leaves[0].transpose(0, 1)).transpose(0, 1))
        else:
            ref = run(lambda d: d.log_prob(leaves[0]))
        for a, b in zip(mine, ref):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-5)
    # operands outside the fast path fall back to torch
    d = Normal(rnd(B, K), rnd(B, K).abs() + 0.1)
    assert _ops.normal_log_prob(d, v) is None and torch.equal(state.log_prob(d, v), d.log_prob(v))


def test_independent_normal_log_prob_fast_path(cuda):
    """state.log_prob on Independent(Normal(loc [B,K,D], scalar scale), 1): equals torch's own log_prob bit for
    bit (one elementwise kernel + the same reduction), gradients match torch autograd; operands outside the fast
    path (per-dimension scales) still go through torch."""
    gen = torch.Generator(device=cuda).manual_seed(0)
    B, K, D = 5, 300, 10
    Normal, Independent = torch.distributions.Normal, torch.distributions.Independent
    for scale in (0.7, torch.tensor(1.3, device=cuda)):
        value = torch.randn(B, K, D, device=cuda, generator=gen).requires_grad_()
        loc = torch.randn(B, K, D, device=cuda, generator=gen).requires_grad_()
        sc = scale.clone().requires_grad_() if torch.is_tensor(scale) else scale
        dist = Independent(Normal(loc, sc, validate_args=False), 1, validate_args=False)
        assert _ops.independent_normal_log_prob(dist, value) is not None
        got = state.log_prob(dist, value)
        want = dist.log_prob(value)
        assert got.shape == (B, K) and torch.equal(got, want)
        cot = torch.randn(B, K, device=cuda, generator=gen)
        leaves = [value, loc] + ([sc] if torch.is_tensor(scale) else [])
        g_got = torch.autograd.grad(got, leaves, cot, retain_graph=True)
        g_want = torch.autograd.grad(want, leaves, cot)
        for a, b in zip(g_got, g_want):
            torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-5)
    # an expanded observation as value (stride 0 over particles) is accepted, per-dimension scales are not
    y = torch.randn(B, D, device=cuda, generator=gen).unsqueeze(1).expand(B, K, D)
    loc = torch.randn(B, K, D, device=cuda, generator=gen)
    dist = Independent(Normal(loc, 0.5, validate_args=False), 1, validate_args=False)
    assert torch.equal(state.log_prob(dist, y), dist.log_prob(y))
    per_dim = Independent(Normal(loc, torch.rand(D, device=cuda, generator=gen) + 0.5, validate_args=False), 1, validate_args=False)
    assert _ops.independent_normal_log_prob(per_dim, y) is None
    assert torch.equal(state.log_prob(per_dim, y), per_dim.log_prob(y))


@pytest.mark.parametrize("K", [4, 64, 1000, 4096, 12000])
def test_warp_per_row_statistics(cuda, K):
    """row_stats.cu (one warp per row, single pass, online max / sum): taken when there are >= 4 rows per SM.
    logsumexp, log-ESS and the weighted moments against float64 references, incl. -inf entries, an all -inf row,
    a +inf row, a NaN row and a row with one dominant particle."""
    B = 1200
    rng = np.random.default_rng(K)
    lw = (rng.standard_normal((B, K)) * rng.uniform(0.3, 15.0, (B, 1))).astype(np.float32)
    lw[1, ::2] = -np.inf
    lw[2] = -np.inf
    lw[3, K // 2] = np.inf
    lw[4, K // 3] = np.nan
    lw[5] = -80.0
    lw[5, K - 1] = 30.0
    lw[6] = 0.25
    x = rng.standard_normal((B, K)).astype(np.float32)
    t, tx = torch.from_numpy(lw).to(cuda), torch.from_numpy(x).to(cuda)
    flags = _ops.new_flags(cuda)
    lse = _ops.logsumexp_rows(t, flags).cpu().numpy()
    ok = np.ones(B, bool)
    ok[[2, 3, 4]] = False
    np.testing.assert_allclose(lse[ok], oracle.lse_f64(lw[ok]), rtol=2e-6, atol=2e-6)
    assert np.isneginf(lse[2]) and np.isposinf(lse[3]) and np.isnan(lse[4])
    assert int(flags.item()) & _lib.FLAG_NAN
    ess = _ops.log_ess_rows(t).cpu().numpy()
    np.testing.assert_allclose(ess[ok], oracle.log_ess_f64(lw[ok]), rtol=2e-5, atol=2e-5)
    assert np.isnan(ess[4])
    mean, second = _ops.weighted_moments(tx, t)
    w = torch.softmax(t[ok].double(), dim=1)
    torch.testing.assert_close(mean[ok, 0].double(), (w * tx[ok].double()).sum(1), rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(second[ok, 0].double(), (w * tx[ok].double() ** 2).sum(1), rtol=2e-5, atol=2e-6)
