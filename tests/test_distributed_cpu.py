"""world_size-2 gloo test of the multi-GPU host logic (row sharding + CSR shard all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import sparse

from graphtools_b200 import distributed as gd


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cpu_scan(x):
    return torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(x.to(torch.int64), 0)])


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = sparse.random(n, n, density=0.01, random_state=1, format="csr", dtype=np.float64)
    bounds = [gd.shard_bounds(n, world, r) for r in range(world)]
    lo, hi = bounds[rank]
    S = M[lo:hi]
    indptr, idx, val = gd.allgather_csr_rows(torch.from_numpy(np.diff(S.indptr).astype(np.int32)),
                                             torch.from_numpy(S.indices.astype(np.int32)),
                                             torch.from_numpy(S.data), [b[1] - b[0] for b in bounds], _cpu_scan)
    ok = (np.array_equal(indptr.numpy(), M.indptr) and np.array_equal(idx.numpy(), M.indices)
          and np.array_equal(val.numpy(), M.data))
    out[rank] = ok
    dist.destroy_process_group()


def test_shard_bounds_cover_rows():
    for n in (1, 127, 128, 1000, 100000, 1000000):
        for world in (1, 2, 4, 8):
            b = [gd.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            for (l0, h0), (l1, h1) in zip(b[:-1], b[1:]):
                assert h0 == l1 and l0 <= h0
            assert all(lo % 128 == 0 or lo == n for lo, _ in b)


def test_allgather_csr_rows_gloo_world2():
    world, n = 2, 700
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _route_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = sparse.random(n, n, density=0.02, random_state=3, format="csr", dtype=np.float64)
    M.setdiag(1.0); M = sparse.csr_matrix(M); M.sort_indices()
    bounds = [gd.shard_bounds(n, world, r) for r in range(world)]
    lo, hi = bounds[rank]
    S = M[lo:hi]
    row_len_t, cols_t, vals_t = gd.route_edges_to_column_owner(
        torch.from_numpy(np.diff(S.indptr).astype(np.int32)), torch.from_numpy(S.indices.astype(np.int32)),
        torch.from_numpy(S.data), lo, bounds)
    T = sparse.csr_matrix(M.T)[lo:hi]          # what this rank must have received: rows lo:hi of M^T
    T.sort_indices()
    ok = (np.array_equal(row_len_t.numpy(), np.diff(T.indptr)) and np.array_equal(cols_t.numpy(), T.indices)
          and np.array_equal(vals_t.numpy(), T.data))
    out[rank] = ok
    dist.destroy_process_group()


def test_route_edges_to_column_owner_gloo_world2():
    world, n = 2, 300
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_route_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_uneven_and_empty_shards_gloo_world2():
    """Query sets that do not fill every rank (out-of-sample extension with few rows): rank 1 owns 44 rows of 300,
    and nothing at all of a 100-row set -- the all-gather and the edge routing must still assemble the full matrix."""
    for n in (300, 100):
        port = _free_port()
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
        assert all(out[r] for r in range(2)), n
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_route_worker, args=(2, port, 100, out), nprocs=2, join=True)
    assert all(out[r] for r in range(2))


def test_owner_of_matches_bounds():
    for n, world in ((1000, 2), (1000, 8), (100, 4), (1_000_000, 8)):
        bounds = [gd.shard_bounds(n, world, r) for r in range(world)]
        cols = torch.arange(0, n, max(1, n // 997))
        own = gd.owner_of(cols, bounds).numpy()
        for c, o in zip(cols.numpy(), own):
            lo, hi = bounds[o]
            assert lo <= c < hi, (n, world, c, o, bounds)
