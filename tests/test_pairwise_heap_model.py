"""A Python model of the tree-free pairwise summation of the multi-CTA path (aesmc_b200/csrc/smc_step_large.cu:
pw_left / pw_leaf_of / pw_node_len, one leaf per multiple-of-64 probe, heap-indexed fold) against numpy's own
float32 np.sum, whose pairwise order scipy.special.logsumexp -- and therefore the reference (math.py:22) -- inherits."""
import numpy as np
import pytest

f32 = np.float32


def pw_left(n):
    h = n >> 1
    return h - (h & 7)


def pw_leaf_of(K, pos):
    start, length, heap = 0, K, 1
    while length > 128:
        n2 = pw_left(length)
        if pos < start + n2:
            length, heap = n2, 2 * heap
        else:
            start, length, heap = start + n2, length - n2, 2 * heap + 1
    return start, length, heap


def pw_node_len(K, h):
    length = K
    for b in range(h.bit_length() - 2, -1, -1):
        if length <= 128:
            return 0
        n2 = pw_left(length)
        length = length - n2 if (h >> b) & 1 else n2
    return length


def leaf_sum(x):
    """numpy's pairwise leaf: 8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail in order."""
    n = len(x)
    if n < 8:
        s = f32(0)
        for v in x:
            s = f32(s + v)
        return s
    lim = n - n % 8
    r = x[:8].copy()
    for i in range(8, lim, 8):
        r = (r + x[i:i + 8]).astype(f32)
    s = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
    for v in x[lim:]:
        s = f32(s + v)
    return s


@pytest.mark.parametrize("K", [129, 1000, 8193, 27000, 100003, 262144])
def test_heap_fold_equals_numpy_sum(K):
    rng = np.random.default_rng(K)
    x = np.exp(rng.standard_normal(K) * 3).astype(f32)
    vals, owners = {}, 0
    for pos in range(0, K, 64):
        start, length, heap = pw_leaf_of(K, pos)
        assert 64 <= length <= 128 or K <= 128
        if pos - 64 < start:                       # the first probe inside the leaf owns it
            assert heap not in vals
            vals[heap] = leaf_sum(x[start:start + length])
            owners += length
    assert owners == K                             # every particle belongs to exactly one owned leaf
    depth = max(h.bit_length() - 1 for h in vals)
    for d in range(depth - 1, -1, -1):
        for h in range(1 << d, 2 << d):
            if pw_node_len(K, h) > 128:
                vals[h] = f32(vals[2 * h] + vals[2 * h + 1])
    assert vals[1].tobytes() == np.sum(x, dtype=f32).tobytes()
