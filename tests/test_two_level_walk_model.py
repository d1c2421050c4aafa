"""A numpy model of the TWO-LEVEL walk of the exact cumulative sum in aesmc_b200/csrc/smc_step_x.cu (round 2), checked
on the CPU against the sequential float32 chain (np.cumsum, inference.py:257).

What the kernel does per row of K = 16 * NT particles (block = a thread's 16 particles, warp = 32 blocks):
  level 1  inside a warp: the parity maps (c0, c1) of consecutive non-mixed blocks compose ON THE BIT PATTERN of the
           chain value (bits + c[bits & 1]); a mixed block cuts the run.  Each warp emits, in particle order, one record
           per mixed block (the map of the run in front of it, then the block) and one for the run that ends its span.
  level 2  one warp, lane = record: the maps between two mixed blocks compose by a segmented scan; only the mixed blocks
           (16 real additions each) are walked serially; the value after every record is scattered to seg_state[warp][k]
           (k = 0: entering the warp's span, k >= 1: after its k-th mixed block).
  replay   every block starts from seg_state[warp][#mixed blocks before it] pushed through the map of the run in front of
           it, and must land on the true chain value.
The records are packed as (c0, c1 - c0 + 1): two chains that start one unit apart stay 0, 1 or 2 units apart.
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
    """Entry value (bit pattern) of every block as the kernel derives it; also returns the record count."""
    blocks, kinds, maps = classify(w)
    nb = len(kinds)
    nw = (nb + 31) // 32
    # ---- level 1 -------------------------------------------------------------------------------------------
    prev_map, kmix, records = [None] * nb, [0] * nb, []
    for wi in range(nw):
        run, k, recs = (0, 0), 0, []
        for b in range(32 * wi, min(32 * wi + 32, nb)):
            prev_map[b], kmix[b] = run, k
            if kinds[b] == 0:
                recs.append((run, b))           # the map of the run in front of the mixed block, then the block
                assert -1 <= run[1] - run[0] <= 1
                run, k = (0, 0), k + 1
            else:
                run = compose(run, maps[b])      # absorbed blocks carry (0, 0)
        if kinds[min(32 * wi + 31, nb - 1)] != 0:
            recs.append((run, -1))
            assert -1 <= run[1] - run[0] <= 1
        records.append(recs)
    # ---- level 2: lane = record, 32 per pass; segmented scan between mixed blocks, mixed blocks serial ---------
    flat = [(wi, r, rec) for wi, recs in enumerate(records) for r, rec in enumerate(recs)]
    seg_state = {(0, 0): 0}
    carry = 0
    for base in range(0, len(flat), 32):
        chunk = flat[base:base + 32]
        comp, head = [], True
        for (_, _, (m, blk)) in chunk:           # inclusive composed map of each record's run of maps
            comp.append(m if head else compose(comp[-1], m))
            head = blk >= 0
        s_own, after_last, carry_in = {}, None, carry
        for i, (_, _, (m, blk)) in enumerate(chunk):
            if blk >= 0:                         # the serial part
                s = from_bits(apply_map(carry, comp[i]))
                carry = bits(seq_sum(blocks[blk], s))
                s_own[i] = carry
        for i, (wi, r, (m, blk)) in enumerate(chunk):
            prior = [j for j in s_own if j < i]
            sg = s_own[max(prior)] if prior else carry_in
            after = s_own[i] if blk >= 0 else apply_map(sg, comp[i])
            last = r == len(records[wi]) - 1
            seg_state[(wi + 1, 0) if last else (wi, r + 1)] = after
            after_last = after
        carry = after_last
    entry = [apply_map(seg_state[(b // 32, kmix[b])], prev_map[b]) for b in range(nb)]
    return entry, kinds, len(flat), carry


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
    # the level-2 warp sees a few dozen records per row, not one per block
    assert nrec / nrows < 8 + 2.5 * (nmixed / nrows) + K / 512
