"""A numpy model of the TWO-LEVEL walk of the exact cumulative sum in aesmc_b200/csrc/smc_step_x.cu (round 2), checked
on the CPU against the sequential float32 chain (np.cumsum, inference.py:257).

What the kernel does per row of K = 16 * NT particles (block = a thread's 16 particles, warp = 32 blocks):
  level 1  inside a warp: the parity maps (c0, c1) of consecutive non-mixed blocks compose ON THE BIT PATTERN of the
           chain value (bits + c[bits & 1]); a mixed block cuts the run.  Every lane ends up with the map of the run in
           front of its block (prev) and of the run through its block (g).
  level 2  a chain over the warps: warp w receives the exact chain value E_w entering its span from warp w - 1 (a tagged
           shared-memory word), and hands E_{w+1} on.  A warp without mixed blocks forwards E_w through the map of its
           32 blocks; otherwise only its mixed blocks (16 real additions each) are walked one after the other.
           seg = chain value after the last mixed block in front of a lane (E_w if there is none).
  replay   every block starts from seg pushed through the map of the run in front of it, and must land on the true
           chain value.
Two chains that start one unit apart stay 0, 1 or 2 units apart (the kernel never stores c1 - c0 any more, the model
still checks it).
"""
import numpy as np
import pytest

from tests.test_exact_scan_model import ITEMS, U24, down, exponent, rows, seq_sum, up

f32 = np.float32


def bits(x):
    return int(np.frombuffer(f32(x).tobytes(), dtype=np.uint32)[0])


def from_bits(b):
    return np.frombuffer(np.uint32(b).tobytes(), dtype=np.float32)[0]


def compose(p, n):
    """(prev then next) on bit patterns: H[q] = P[q] + N[(q + P[q]) & 1]"""
    return (p[0] + n[p[0] & 1], p[1] + n[(1 + p[1]) & 1])


def apply_map(b, m):
    return b + m[b & 1]


def classify(w):
    """Per block: kind (> 0 pure binade, 0 mixed, -1 absorbed) and parity map, as the kernel's P3 does."""
    nb = len(w) // ITEMS
    blocks = w.reshape(nb, ITEMS)
    block_sum = np.array([seq_sum(r) for r in blocks], f32)
    prefix = np.concatenate([[f32(0)], np.cumsum(block_sum, dtype=f32)])
    kinds, maps = [], []
    for b in range(nb):
        eps = f32(ITEMS * (b + 1) + 64) * U24
        lo, hi = down(f32(prefix[b]) * f32(1 - eps)), up(f32(prefix[b + 1]) * f32(1 + eps))
        eb = 0
        if lo >= f32(2.0 ** -100) and exponent(lo) == exponent(hi):
            eb = exponent(lo)
        if block_sum[b] == 0 or block_sum[b] < down(lo * f32(2.0 ** -26)):
            eb = -1
        if eb == -1 and b % 32:
            live = [j for j in range(b - b % 32, b) if kinds[j] != -1]
            if live and kinds[live[-1]] > 0:
                eb = kinds[live[-1]]
        c = (0, 0)
        if eb > 0:
            scale = f32(2.0 ** (23 - (eb - 127)))
            m0, m1 = f32(2 ** 23), f32(2 ** 23 + 1)
            for v in blocks[b]:
                sv = f32(v * scale)
                m0, m1 = f32(m0 + sv), f32(m1 + sv)
            if m1 < f32(2 ** 24):
                c = (int(m0) - 2 ** 23, int(m1) - 2 ** 23 - 1)
            else:
                eb = 0
        kinds.append(eb)
        maps.append(c)
    return blocks, kinds, maps


def two_level_entry_states(w):
    """Entry value (bit pattern) of every block as the kernel derives it; also returns the number of serial steps."""
    blocks, kinds, maps = classify(w)
    nb = len(kinds)
    nw = (nb + 31) // 32
    entry, steps = [0] * nb, 0
    E = 0                                        # bits of the chain value entering warp 0's span: 0.0f
    for wi in range(nw):
        lanes = range(32 * wi, min(32 * wi + 32, nb))
        # ---- level 1: per lane, the map of the run in front of the block (prev) and through it (g) -------------
        run, prev, g = (0, 0), {}, {}
        for b in lanes:
            prev[b] = run
            assert -1 <= run[1] - run[0] <= 1
            run = (0, 0) if kinds[b] == 0 else compose(run, maps[b])   # absorbed blocks carry (0, 0)
            g[b] = run
        # ---- level 2: E_w -> seg of every lane -> E_{w+1} --------------------------------------------------------
        seg = {b: E for b in lanes}
        carry = E
        for b in lanes:
            if kinds[b] == 0:                    # the serial part: one step per mixed block
                s = from_bits(apply_map(carry, prev[b]))
                carry = bits(seq_sum(blocks[b], s))
                steps += 1
                for later in lanes:
                    if later > b:
                        seg[later] = carry
        last = lanes[-1]
        for b in lanes:
            entry[b] = apply_map(seg[b], prev[b])
        E = carry if kinds[last] == 0 else apply_map(seg[last], g[last])
        steps += 1                               # the hand-off to the next warp
    return entry, kinds, steps, E


@pytest.mark.parametrize("K", [1024, 4096, 16384])
def test_two_level_walk_reproduces_the_sequential_chain(K):
    rng = np.random.default_rng(K)
    nrec, nmixed, nrows = 0, 0, 0
    for name, lw in rows(K, rng):
        lw = lw.astype(f32)
        with np.errstate(over="ignore", invalid="ignore"):
            e = np.exp((lw - lw.max()).astype(f32)).astype(f32)
            w = (e / e.sum(dtype=f32)).astype(f32)
        chain = np.cumsum(w, dtype=f32)
        entry, kinds, n, total = two_level_entry_states(w)
        truth = np.concatenate([[f32(0)], chain[ITEMS - 1::ITEMS]])
        got = np.array([from_bits(b) for b in entry], f32)
        # every pure block's binade assumption holds at its exact entry (what the kernel re-verifies), and every
        # block's entry value is the chain's, bit for bit
        assert np.array_equal(got.view(np.uint32), truth[:-1].view(np.uint32)), name
        assert total == bits(chain[-1]), name
        for b, k in enumerate(kinds):
            if k > 0:
                assert exponent(got[b]) == k or got[b] == 0, (name, b)
        nrec += n
        nmixed += sum(k == 0 for k in kinds)
        nrows += 1
    # the serial part of a row is one step per mixed block and one hand-off per warp, not one per block
    assert nrec / nrows <= nmixed / nrows + K / 512
