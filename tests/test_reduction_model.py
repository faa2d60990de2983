"""Executable model of the halving-exchange warp reduction in the float64 refine (csrc/refine.cu:
GTB_REDUCE8_AND_STORE): eight per-lane partial sums are reduced with 4 + 2 + 1 exchanges and two butterfly steps.
The kernel's claim is that every result is bit-identical to the plain xor-butterfly used by warp_dist2 (so both
refine stages return the same bits); the model checks exactly that on random float64 data."""
import numpy as np


def butterfly(vals):
    """vals[lane] -> the value every lane holds after xor-reductions with offsets 16, 8, 4, 2, 1."""
    v = np.array(vals, dtype=np.float64)
    for off in (16, 8, 4, 2, 1):
        v = v + v[np.arange(32) ^ off]
    return v


def halving(acc):
    """acc[lane][u] (8 accumulators per lane) -> (value per lane, candidate index per lane)."""
    lanes = np.arange(32)
    acc = np.array(acc, dtype=np.float64)                  # [32, 8]
    up = (lanes & 16) != 0
    keep = np.where(up[:, None], acc[:, 4:8], acc[:, 0:4])
    give = np.where(up[:, None], acc[:, 0:4], acc[:, 4:8])
    h4 = keep + give[lanes ^ 16]
    up = (lanes & 8) != 0
    keep = np.where(up[:, None], h4[:, 2:4], h4[:, 0:2])
    give = np.where(up[:, None], h4[:, 0:2], h4[:, 2:4])
    h2 = keep + give[lanes ^ 8]
    up = (lanes & 4) != 0
    keep = np.where(up, h2[:, 1], h2[:, 0])
    give = np.where(up, h2[:, 0], h2[:, 1])
    h1 = keep + give[lanes ^ 4]
    h1 = h1 + h1[lanes ^ 2]
    h1 = h1 + h1[lanes ^ 1]
    myu = ((lanes >> 4) & 1) * 4 + ((lanes >> 3) & 1) * 2 + ((lanes >> 2) & 1)
    return h1, myu


def test_halving_exchange_equals_butterfly_bit_for_bit():
    rng = np.random.default_rng(0)
    for trial in range(200):
        scale = 10.0 ** rng.integers(-8, 9)
        acc = rng.standard_normal((32, 8)) ** 2 * scale    # partial sums of squares, widely varying magnitude
        h1, myu = halving(acc)
        for u in range(8):
            ref = butterfly(acc[:, u])
            got = h1[myu == u]
            assert len(got) == 4
            assert np.all(got.view(np.uint64) == ref[myu == u].view(np.uint64)), (trial, u)


def test_every_candidate_lands_on_a_writer_lane():
    _, myu = halving(np.zeros((32, 8)))
    writers = [l for l in range(32) if l % 4 == 0]
    assert sorted(myu[writers]) == list(range(8))
