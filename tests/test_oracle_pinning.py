"""The oracle is only worth something if it is pinned to the reference (CPU-only tests).

1. every restated third-party primitive (numpy float32 exp/log/pairwise-sum, glibc log1pf, scipy
   logsumexp) is compared bit-for-bit with the library installed in this image;
2. the C restatement and the torch/numpy port are compared with tests/golden/*.npz, recorded by
   tests/golden/make_golden.py from the UNMODIFIED reference imported from /root/reference;
3. the reference's own known-answer vectors (test/test_state.py:286-303, test/test_inference.py:13-40,
   test/test_math.py:51-64,111-126, test/test_statistics.py:32-42,71-115) are replayed on the port.
"""
import ctypes
import ctypes.util

import numpy as np
import pytest
import torch

from oracle import core, kalman, reference_port as port
from tests.models import lgssm


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def vec(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    f = getattr(core.lib(), fn)
    for i, v in enumerate(x):
        out[i] = f(ctypes.c_float(float(v)))
    return out


# ---- 1. primitives vs the installed numpy / glibc / scipy -----------------------------------------
def test_np_exp_bits():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-104.5, 0, 60000), rng.uniform(0, 89, 5000), rng.uniform(-1e-3, 1e-3, 2000),
                        [-np.inf, 0.0, -0.0, -103.97208, -103.9720841, 88.7228394, -87.3, -88.0, -100.0]]).astype(np.float32)
    with np.errstate(over="ignore"):
        assert np.array_equal(bits(vec("aesmc_oracle_np_expf", x)), bits(np.exp(x)))


def test_np_log_bits():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.arange(1, 5000), 2.0 ** np.arange(0, 24), np.exp(rng.uniform(-80, 80, 30000))]).astype(np.float32)
    assert np.array_equal(bits(vec("aesmc_oracle_np_logf", x)), bits(np.log(x)))


def test_log1pf_bits_vs_glibc():
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.log1pf.restype = ctypes.c_float
    libm.log1pf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(0, 1, 20000), rng.uniform(0, 70000, 20000), np.exp(rng.uniform(-40, 16, 20000)),
                        [0.0, 1.0, 2.0 ** -30, 0.41421357]]).astype(np.float32)
    ref = np.array([libm.log1pf(ctypes.c_float(float(v))) for v in x], np.float32)
    assert np.array_equal(bits(vec("aesmc_oracle_log1pf", x)), bits(ref))


@pytest.mark.parametrize("K", [1, 2, 7, 8, 9, 127, 128, 129, 130, 1000, 4095, 4096, 4097, 8193, 65536, 100003])
def test_pairwise_sum_bits(K):
    rng = np.random.default_rng(K)
    a = rng.random((5, K), dtype=np.float32)
    assert np.array_equal(bits(core.pairwise_sum_rows(a)), bits(np.sum(a, axis=1)))


def test_port_logsumexp_is_scipy():
    scipy_special = pytest.importorskip("scipy.special")
    rng = np.random.default_rng(3)
    for K in (1, 2, 3, 100, 4096):
        a = (rng.standard_normal((32, K)) * 3).astype(np.float32)
        a[0, 0] = -np.inf
        with np.errstate(all="ignore"):
            assert np.array_equal(bits(port.np_logsumexp_rows(a)), bits(scipy_special.logsumexp(a, axis=1, keepdims=True)))


# ---- 2. golden vectors from the unmodified reference ----------------------------------------------
def _step_names(g):
    return sorted({k.split("/")[1] for k in g.files if k.startswith("step/")})


@pytest.mark.parametrize("variant", ["avx2", "default"])
def test_c_oracle_reproduces_reference_indices(golden, variant):
    g = golden[variant]
    total = 0
    for name in _step_names(g):
        lw, u, idx, lse = (g["step/%s/%s" % (name, k)] for k in ("lw", "u", "idx", "lse"))
        mine, status, my_lse, w, _ = core.sample_ancestral_index(lw, u, return_parts=True)
        assert status == core.OK
        assert np.array_equal(mine, idx.astype(np.int64)), name
        if variant == "avx2":  # libm log1p: every intermediate is pinned too
            assert np.array_equal(bits(my_lse), bits(lse)), name
            if "step/%s/w" % name in g.files:
                assert np.array_equal(bits(w), bits(g["step/%s/w" % name])), name
        else:  # SVML log1p rows replayed through the injection port
            mine2, _ = core.sample_ancestral_index(lw, u, lse_inject=lse)
            assert np.array_equal(mine2, idx.astype(np.int64)), name
        total += idx.size
    assert total > 150000


def _build_models(params, proposal_state):
    m0, s0, a, sx, c, sy, q0, _ = params
    init = lgssm.Initial(float(m0), float(s0))
    trans = lgssm.Transition(a, float(sx))
    emis = lgssm.Emission(c, float(sy))
    prop = lgssm.Proposal(float(q0), float(q0))  # the reference model uses scale_0 at every step
    flat = torch.from_numpy(np.asarray(proposal_state, dtype=np.float32))
    off = 0
    for p in prop.parameters():
        p.data.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    return init, trans, emis, prop


@pytest.mark.parametrize("tag,algo", [("c1_smc", "smc"), ("small_smc", "smc"), ("small_is", "is"), ("t1_smc", "smc")])
def test_port_reproduces_reference_infer(golden, tag, algo):
    g = golden["default"]
    p = "infer/%s/" % tag
    obs = [torch.from_numpy(o) for o in g[p + "obs"]]
    seed = int(g[p + "seed"][0])
    models = _build_models(g[p + "params"], g[p + "proposal_state"])
    smc = algo == "smc"
    u = g[p + "u"]
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        res = port.infer(algo, obs, *models, g[p + "log_weights"].shape[2], return_log_marginal_likelihood=True,
                         return_latents=True, return_original_latents=smc, return_log_weight=True,
                         return_log_weights=True, return_ancestral_indices=smc,
                         uniforms=[u[t] for t in range(u.shape[0])] if smc else None)
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.stack(res["log_weights"]).numpy(), g[p + "log_weights"], **tol)
    np.testing.assert_allclose(res["log_marginal_likelihood"].numpy(), g[p + "lml"], **tol)
    np.testing.assert_allclose(res["log_weight"].numpy(), g[p + "log_weight"], **tol)
    np.testing.assert_allclose(torch.stack(res["latents"]).numpy(), g[p + "latents"], **tol)
    if smc:
        got = (torch.stack(res["ancestral_indices"]).numpy() if res["ancestral_indices"]
               else np.zeros_like(g[p + "ancestral_indices"]))
        assert np.array_equal(got, g[p + "ancestral_indices"])
        np.testing.assert_allclose(torch.stack(res["original_latents"]).numpy(), g[p + "original_latents"], **tol)
    # losses.get_loss on the same seeds
    np.random.seed(seed)
    torch.manual_seed(seed + 1)
    loss = port.get_loss(obs, g[p + "log_weights"].shape[2], "aesmc" if smc else "iwae", *models)
    np.testing.assert_allclose(loss.item(), g[p + "loss"][0], rtol=1e-5)


def test_port_statistics_match_reference(golden):
    g = golden["default"]
    lw = torch.from_numpy(g["stats/lw"])
    val = torch.from_numpy(g["stats/value"])
    np.testing.assert_allclose(port.log_ess(lw).numpy(), g["stats/log_ess"], rtol=1e-6)
    np.testing.assert_allclose(port.weighted_expectation(val, lw, lambda x: x).numpy(), g["stats/mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(core.log_ess_f64(g["stats/lw"]), g["stats/log_ess"], rtol=1e-5)


# ---- 3. the reference's own known-answer vectors --------------------------------------------------
def test_resample_known_answer():
    # test/test_state.py:286-303
    value = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    idx = np.array([[1, 2, 0], [0, 0, 1]])
    want = np.array([[2, 3, 1], [4, 4, 5]], np.float32)
    assert np.array_equal(core.resample(value, idx), want)
    assert torch.equal(port.gather_particles(torch.from_numpy(value), torch.from_numpy(idx)), torch.from_numpy(want))


def test_genealogy_known_answer():
    # test/test_inference.py:13-40
    latents = [torch.Tensor([[1, 2, 3]]), torch.Tensor([[4, 5, 6]]), torch.Tensor([[7, 8, 9]]), torch.Tensor([[10, 11, 12]])]
    anc = [torch.LongTensor([[0, 2, 1]]), torch.LongTensor([[2, 0, 0]]), torch.LongTensor([[1, 2, 0]])]
    want = [[1, 1, 2], [4, 4, 6], [8, 9, 7], [10, 11, 12]]
    got = port.trace_genealogy(latents, anc)
    for g, w in zip(got, want):
        assert g[0].tolist() == w
    # the C composition primitive gives the same cursor
    cur = np.arange(3)[None]
    for t in (2, 1, 0):
        cur = core.compose_index(anc[t].numpy(), cur)
    assert cur.tolist() == [[0, 0, 1]]


def test_softmax_known_answer():
    # test/test_math.py:51-64, 111-126
    x = np.array([[1.0, 2.0, 3.0]], np.float32)
    w, lse = core.normalized_weights(x)
    want = np.exp(x) / np.exp(x).sum()
    np.testing.assert_allclose(w, want, rtol=1e-6)
    np.testing.assert_allclose(x - lse[:, None], np.log(want), atol=1e-6)


def test_log_ess_known_answer():
    # test/test_statistics.py:71-115 (float64, +-1e6 offsets)
    nw = np.array([0.2, 0.3, 0.5])
    for shift in (np.log(0.47), 1e6, -1e6):
        lw = torch.from_numpy(np.log(nw) + shift)
        assert abs(port.log_ess(lw).item() - np.log(1 / np.sum(nw ** 2))) < 1e-7


def test_sampler_frequencies():
    # test/test_inference.py:64-84: 10 000 rows of weights [.2,.3,.5]
    rng = np.random.default_rng(0)
    lw = np.log(np.tile(np.array([[0.2, 0.3, 0.5]], np.float32), (10000, 1)))
    idx, st = core.sample_ancestral_index(lw, rng.random(10000))
    freq = np.bincount(idx.ravel(), minlength=3) / idx.size
    np.testing.assert_allclose(freq, [0.2, 0.3, 0.5], atol=1e-2)


def test_port_smc_evidence_tracks_kalman():
    """Bootstrap filter log-evidence vs the exact Kalman value (O(sqrt(T/K)) per row)."""
    T, B, K = 30, 6, 2000
    ys = lgssm.simulate(T, B, seed=3)
    models = lgssm.bootstrap_filter()
    torch.manual_seed(0)
    np.random.seed(0)
    with torch.no_grad():
        res = port.infer("smc", [torch.from_numpy(y) for y in ys], *models, K,
                         return_log_marginal_likelihood=True, return_latents=False)
    exact = kalman.lgssm1d_log_evidence(ys, 0.0, 1.0, 0.9, 1.0, 1.0, 0.25)
    assert np.max(np.abs(res["log_marginal_likelihood"].numpy() - exact)) < 0.8
