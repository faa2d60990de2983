"""Executable model of the warp-cooperative quickselect that compacts a candidate buffer in the search epilogue
(csrc/search_tc.cu: compact_row): the same slot layout (element e -> slot e // 32, lane e % 32), pivot rotation,
interval narrowing and survivor selection, written with Python integers standing in for ballots.  It pins the two
properties the kernel relies on -- the survivors are exactly the LS smallest keys, and the loop terminates within
the number of buffered elements whatever their arrival order -- and documents the expected number of pivot steps."""
import numpy as np
import pytest

NONE = (1 << 64) - 1


def f2ord(bits):
    """order-preserving map of float32 bit patterns to uint32 (search_tc.cu: f2ord)."""
    return bits ^ (0xFFFFFFFF if bits & 0x80000000 else 0x80000000)


def compact_model(keys, ls, nslot=3):
    slots = [[NONE] * 32 for _ in range(nslot)]
    for e, k in enumerate(keys):
        slots[e // 32][e % 32] = k
    lo, hi, T, steps = 0, NONE, NONE, 0
    for it in range(nslot * 32 + 1):
        steps += 1
        masks = [sum(1 << l for l in range(32) if lo < slots[i][l] < hi) for i in range(nslot)]
        cand, mm = None, 0
        for i in range(nslot):
            sl = (i + it) % nslot
            if mm == 0 and masks[sl]:
                cand, mm = slots[sl], masks[sl]
        if mm == 0:
            break
        src = (mm.bit_length() - 1) if (it & 2) else ((mm & -mm).bit_length() - 1)
        pv = cand[src]
        c = sum(1 for i in range(nslot) for l in range(32) if slots[i][l] < pv)
        if c == ls - 1:
            T = pv
            break
        if c >= ls:
            hi = pv
        else:
            lo = pv
    kept = sorted(k for row in slots for k in row if k <= T)
    return kept, steps


@pytest.mark.parametrize("ls", [16, 32])
@pytest.mark.parametrize("order", ["random", "ascending", "descending"])
def test_survivors_are_the_ls_smallest(ls, order):
    rng = np.random.default_rng(ls + len(order))
    steps = []
    for _ in range(300):
        cnt = int(rng.integers(ls + 1, 97))
        vals = rng.standard_normal(cnt).astype(np.float32) * 50.0          # negative and positive values
        idx = rng.choice(1 << 20, size=cnt, replace=False)
        keys = [(f2ord(int(v.view(np.uint32))) << 32) | int(i) for v, i in zip(vals, idx)]
        if order == "ascending":
            keys.sort()
        elif order == "descending":
            keys.sort(reverse=True)
        kept, n = compact_model(keys, ls)
        assert kept == sorted(keys)[:ls]
        assert n <= cnt
        steps.append(n)
    assert np.mean(steps) < (12 if order == "random" else 40)


def test_ordered_key_is_monotone_in_the_float_value():
    vals = np.array([-np.inf, -3.5e30, -1.0, -1e-30, -0.0, 0.0, 1e-30, 2.0, 7.5e29, np.inf], dtype=np.float32)
    ords = [f2ord(int(v.view(np.uint32))) for v in vals]
    assert ords == sorted(ords) and len(set(ords)) == len(ords)


def test_ties_in_value_are_broken_by_index():
    v = f2ord(int(np.float32(1.25).view(np.uint32)))
    keys = [(v << 32) | i for i in (9, 3, 7, 1, 5)] + [((v + 1) << 32) | 0]
    kept, _ = compact_model(keys + [((v + 2) << 32) | j for j in range(40)], 4)
    assert kept == [(v << 32) | i for i in (1, 3, 5, 7)]
