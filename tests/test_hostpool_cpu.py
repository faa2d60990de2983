"""Host-side bookkeeping of the page-locked result pool (graphtools_b200/hostpool.py) on CPU: layout, hand-out,
release when the last view dies, reuse, and that scipy keeps the handed-out arrays instead of copying them (its
constructors copy any array that is a view of a much larger ndarray).  Pinning itself needs CUDA and is covered by
tests/test_parity_gpu.py::test_host_result_pool_recycles_without_aliasing."""
import gc

import numpy as np
import pytest
import torch
from scipy import sparse

from graphtools_b200 import hostpool as hp


@pytest.fixture
def unpinned(monkeypatch):
    real = torch.empty

    def empty(*a, **k):
        k.pop("pin_memory", None)
        return real(*a, **k)
    monkeypatch.setattr(torch, "empty", empty)
    monkeypatch.setattr(hp, "_local", [])
    yield


def test_layout_is_aligned_and_disjoint():
    lay, total = hp.layout([("a", 10, np.int32), ("b", 5, np.float64), ("c", 0, np.float64), ("d", 3, np.int32)])
    offs = [lay[k][0] for k in "abcd"]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and total % 256 == 0
    assert total >= offs[-1] + 12


def test_block_is_busy_until_the_last_view_dies_and_then_reused(unpinned):
    lay, total = hp.layout([("vals", 2_000_000, np.float64), ("indices", 2_000_000, np.int32), ("indptr", 1001, np.int32)])
    blk = hp.take_local(total)
    assert blk is not None and blk.free
    out = blk.carve(lay)
    assert not blk.free
    out["indptr"][:] = np.linspace(0, 2_000_000, 1001).astype(np.int32)
    out["indices"][:] = 0
    out["vals"][:] = 1.0
    M = sparse.csr_matrix((out["vals"], out["indices"], out["indptr"]), shape=(1000, 5_000_000))
    assert np.shares_memory(M.data, out["vals"]) and np.shares_memory(M.indices, out["indices"])
    other = hp.take_local(total)                      # the first block is busy: a second one
    assert other is not blk
    del out
    gc.collect()
    assert not blk.free                                # M still holds the arrays
    row = M.indices
    del M
    gc.collect()
    assert not blk.free                                # one view left
    del row
    gc.collect()
    assert blk.free
    assert hp.take_local(total) is blk                 # recycled


def test_pool_respects_the_cap_and_the_switch(unpinned, monkeypatch):
    monkeypatch.setenv("GTB_HOST_POOL_MB", "8")
    assert hp.take_local(16 << 20) is None             # larger than the cap
    small = hp.take_local(1 << 20)
    assert small is not None
    monkeypatch.setenv("GTB_HOST_POOL", "0")
    assert hp.take_local(1 << 20) is None
