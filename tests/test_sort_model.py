"""Executable model of warp_sort32 (csrc/common.cuh): the 15-step bitonic network over one (key, value) pair per
lane with the kernel's exact take rule, including ties in the key (broken by the value) and fully equal pairs
(padding), which must not be exchanged inconsistently."""
import numpy as np


def warp_sort32_model(k, v):
    k, v = np.array(k), np.array(v)
    lanes = np.arange(32)
    kk = 2
    while kk <= 32:
        j = kk >> 1
        while j > 0:
            ok, ov = k[lanes ^ j], v[lanes ^ j]
            keep_min = ((lanes & kk) == 0) == ((lanes & j) == 0)
            gt = (k > ok) | ((k == ok) & (v > ov))
            lt = (k < ok) | ((k == ok) & (v < ov))
            take = np.where(keep_min, gt, lt)
            k, v = np.where(take, ok, k), np.where(take, ov, v)
            j >>= 1
        kk <<= 1
    return k, v


def test_sorts_by_key_then_value_and_keeps_the_multiset():
    rng = np.random.default_rng(1)
    for trial in range(300):
        n_valid = int(rng.integers(0, 33))
        k = np.full(32, np.inf)
        v = np.full(32, 0x7FFFFFFF, dtype=np.int64)
        k[:n_valid] = rng.integers(0, 6, size=n_valid).astype(np.float64)     # many ties in the key
        v[:n_valid] = rng.permutation(1000)[:n_valid]
        perm = rng.permutation(32)
        ks, vs = warp_sort32_model(k[perm], v[perm])
        order = np.lexsort((v, k))
        assert np.array_equal(ks, k[order]) and np.array_equal(vs, v[order]), trial


def test_column_sort_with_padding_keeps_payloads_attached():
    rng = np.random.default_rng(2)
    cols = np.full(32, 0x7FFFFFFF, dtype=np.int64)
    w = np.zeros(32)
    cols[:11] = rng.permutation(5000)[:11]
    w[:11] = rng.random(11)
    perm = rng.permutation(32)
    cs, ws = warp_sort32_model(cols[perm], w[perm])
    assert np.array_equal(cs[:11], np.sort(cols[:11]))
    lookup = dict(zip(cols[:11], w[:11]))
    assert all(ws[i] == lookup[cs[i]] for i in range(11))
