"""A numpy model of aesmc_b200/csrc/exact_scan.cuh -- the parallel evaluation of np.cumsum's sequential float32
chain -- checked on the CPU against the sequential chain itself.

The CUDA code cannot run here; what can be checked without a GPU is the arithmetic it rests on:
  * inside one binade a block of additions is the map m -> m + c[m & 1] on the integer mantissa, with (c0, c1) read
    off two scaled float chains started at 2^23 and 2^23 + 1;
  * the block classification from an approximate prefix with the k * 2^-24 error bound never calls a block "pure"
    that is not;
  * a block whose weights sum to less than 2^-26 of a lower bound of its entry value leaves the chain untouched
    ("absorbed": identity in every binade) and may join the pure run it follows.
Every state the model hands from block to block must equal the true sequential value bit for bit; on the GPU the same
invariants are re-verified at run time and the stress tests in test_step_parity_gpu.py cover the kernels themselves.
"""
import numpy as np
import pytest

f32 = np.float32
ITEMS = 16
U24 = f32(2.0 ** -24)


def down(x):
    return np.nextafter(f32(x), f32(-np.inf))


def up(x):
    return np.nextafter(f32(x), f32(np.inf))


def exponent(x):
    return int(np.frombuffer(f32(x).tobytes(), dtype=np.uint32)[0] >> 23)


def seq_sum(values, start=f32(0)):
    s = f32(start)
    for v in values:
        s = f32(s + v)
    return s


def model_block_states(w):
    """Entry state of every 16-particle block as exact_scan.cuh would derive it (maps for pure blocks, sequential
    additions for mixed ones, nothing for absorbed ones); returns (states, kinds, failed)."""
    K = len(w)
    nb = (K + ITEMS - 1) // ITEMS
    w = np.concatenate([w, np.zeros(nb * ITEMS - K, f32)]).reshape(nb, ITEMS)
    block_sum = np.array([seq_sum(row) for row in w], f32)
    prefix = np.concatenate([[f32(0)], np.cumsum(block_sum, dtype=f32)])  # any float order: the bound covers it
    kinds, maps = [], []
    for b in range(nb):
        p_in, p_out = prefix[b], prefix[b + 1]
        eps = f32((ITEMS * (b + 1) + 64)) * U24
        lo = down(f32(p_in) * f32(1 - eps))
        hi = up(f32(p_out) * f32(1 + eps))
        eb = 0
        if lo >= f32(2.0 ** -100) and exponent(lo) == exponent(hi):
            eb = exponent(lo)
        if block_sum[b] == 0 or block_sum[b] < down(lo * f32(2.0 ** -26)):
            eb = -1
        if eb == -1 and b % 32:  # joins the pure run it follows (same warp of 32 blocks)
            live = [j for j in range(b - b % 32, b) if kinds[j] != -1]
            if live and kinds[live[-1]] > 0:
                eb = kinds[live[-1]]
        c = (0, 0)
        if eb > 0:
            scale = f32(2.0 ** (23 - (eb - 127)))
            m0, m1 = f32(2 ** 23), f32(2 ** 23 + 1)
            for v in w[b]:
                sv = f32(v * scale)
                m0, m1 = f32(m0 + sv), f32(m1 + sv)
            if m1 < f32(2 ** 24):
                c = (int(m0) - 2 ** 23, int(m1) - 2 ** 23 - 1)
            else:
                eb = 0
        kinds.append(eb)
        maps.append(c)
    states, s, failed = [], f32(0), False
    for b in range(nb):
        states.append(s)
        eb = kinds[b]
        if eb > 0:
            if exponent(s) != eb:
                failed = True
                break
            bits = int(np.frombuffer(f32(s).tobytes(), dtype=np.uint32)[0])
            m = (bits & 0x7FFFFF) | 0x800000
            m += maps[b][m & 1]
            if m > 0x1000000:
                failed = True
                break
            s = f32(2.0 ** (eb - 127 + 1)) if m == 0x1000000 else f32(m * 2.0 ** (eb - 127 - 23))
        elif eb == 0:
            s = seq_sum(w[b], s)
    return states, kinds, failed


def rows(K, rng):
    yield "typical", rng.standard_normal(K) - 1.4
    yield "heavy tailed", rng.standard_normal(K) * 6
    yield "collapsed", rng.standard_normal(K) * 20
    yield "increasing", np.sort(rng.standard_normal(K) * 3)
    yield "decreasing", np.sort(rng.standard_normal(K) * 3)[::-1]
    yield "equal", np.full(K, -0.5)
    yield "dyadic", np.log(2.0) * rng.integers(-30, 0, K)
    yield "mostly zero", np.where(rng.random(K) < 0.9, -np.inf, rng.standard_normal(K))
    dominated = rng.standard_normal(K) * 0.01 - 70.0
    dominated[K // 3] = 0.0
    yield "one dominant particle", dominated
    yield "long sub-ulp prefix", np.concatenate([np.full(K // 2, -40.0), rng.standard_normal(K - K // 2)])
    for i in range(24):  # random spreads from benign to fully collapsed, some with exact ties (quantised weights)
        lw = rng.standard_normal(K) * rng.uniform(0.2, 25)
        yield "random spread %d" % i, (np.round(lw * 4) / 4 if i % 3 == 0 else lw)


@pytest.mark.parametrize("K", [100, 1000, 4096])
def test_block_states_equal_the_sequential_chain(K):
    rng = np.random.default_rng(K)
    absorbed = pure = mixed = 0
    for name, lw in rows(K, rng):
        lw = lw.astype(f32)
        with np.errstate(over="ignore", invalid="ignore"):
            m = lw.max()
            e = np.exp((lw - m).astype(f32)).astype(f32)
            w = (e / e.sum(dtype=f32)).astype(f32)  # any normalisation will do for this test
        chain = np.cumsum(w, dtype=f32)  # numpy's float32 cumsum is the sequential chain (oracle pinning test)
        states, kinds, failed = model_block_states(w)
        assert not failed, name
        truth = np.concatenate([[f32(0)], chain[ITEMS - 1::ITEMS]])[: len(states)]
        got = np.array(states, f32)
        assert np.array_equal(got.view(np.uint32), truth.view(np.uint32)), name
        absorbed += sum(k == -1 for k in kinds)
        pure += sum(k > 0 for k in kinds)
        mixed += sum(k == 0 for k in kinds)
    if K >= 1000:
        assert pure > mixed   # the classification is not vacuous: most blocks are handled by maps
        assert absorbed > 0   # and the sparse / dominated rows exercise the absorbed rule


def test_parity_map_is_the_rounding_rule():
    """m -> m + c[m & 1] reproduces 16 float additions for EVERY mantissa of a binade sample, both parities."""
    rng = np.random.default_rng(0)
    for trial in range(200):
        e = int(rng.integers(-60, 1))
        u = 2.0 ** (e - 23)
        w = (rng.random(ITEMS) * rng.choice([0.3, 3.0, 40.0]) * u).astype(f32)  # sub-ulp ... tens of ulps, ties possible
        if trial % 5 == 0:
            w = (np.round(w / f32(u / 2)) * f32(u / 2)).astype(f32)             # exact half-ulp multiples: ties
        scale = f32(2.0 ** (23 - e))
        m0, m1 = f32(2 ** 23), f32(2 ** 23 + 1)
        for v in w:
            m0, m1 = f32(m0 + f32(v * scale)), f32(m1 + f32(v * scale))
        c = (int(m0) - 2 ** 23, int(m1) - 2 ** 23 - 1)
        for m in rng.integers(2 ** 23, 2 ** 24 - 1 - max(c) - 1, 50):
            m = int(m)
            s = seq_sum(w, f32(m * u))
            assert s == f32((m + c[m & 1]) * u), (trial, m)
