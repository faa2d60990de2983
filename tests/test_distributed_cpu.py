"""world_size-2 gloo tests of the multi-GPU host logic: row sharding, CSR shard all-gather, the edge exchange
(split-size exchange + one all-to-all of packed records) and the shared-memory result assembly."""
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


def _torch_bucket(indptr, indices, data, lo, per, world):
    """torch restatement of csrc/symm.cu route_count / route_fill (test helper: the product path buckets with the
    CUDA kernels): records {int32 i, int32 j, float64 w} in (destination, row, column) order + per-rank counts."""
    m = indptr.shape[0] - 1
    rows = torch.repeat_interleave(torch.arange(lo, lo + m, dtype=torch.int64), indptr[1:] - indptr[:-1])
    dest = torch.clamp(indices.to(torch.int64) // per, max=world - 1)
    order = torch.argsort(dest, stable=True)
    ij = torch.stack([rows.to(torch.int32), indices.to(torch.int32)], 1)[order].contiguous().view(torch.int64).view(-1)
    w = data[order].contiguous().view(torch.int64)
    return torch.stack([ij, w], 1).contiguous(), torch.bincount(dest, minlength=world)


def _route_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = sparse.random(n, n, density=0.02, random_state=3, format="csr", dtype=np.float64)
    M.setdiag(1.0); M = sparse.csr_matrix(M); M.sort_indices()
    bounds = [gd.shard_bounds(n, world, r) for r in range(world)]
    lo, hi = bounds[rank]
    S = M[lo:hi]
    rec = gd.exchange_edges(torch.from_numpy(S.indptr.astype(np.int64)), torch.from_numpy(S.indices.astype(np.int32)),
                            torch.from_numpy(S.data), lo, bounds, bucket_fn=_torch_bucket)
    i, j, w = (t.numpy() for t in gd.unpack_records(rec))
    # what this rank must have received: the entries of rows lo:hi of M^T, i.e. (j - lo, i) -> w
    got = sparse.csr_matrix((w, (j - lo, i)), shape=(hi - lo, n)) if hi > lo else None
    T = sparse.csr_matrix(M.T)[lo:hi]
    T.sort_indices()
    ok = len(i) == T.nnz and ((j >= lo) & (j < hi)).all()
    if hi > lo:
        got.sort_indices()
        ok = ok and (np.array_equal(got.indptr, T.indptr) and np.array_equal(got.indices, T.indices)
                     and np.array_equal(got.data, T.data))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _shared_worker(rank, world, port, out):
    """SharedResult: every rank fills its slice of a shared-memory array; all ranks then see the whole array."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = gd.SharedResult()
    arr = res.create({"a": (1000, np.float64), "b": (10, np.int32)})
    per = 1000 // world
    arr["a"][rank * per:(rank + 1) * per] = np.arange(rank * per, (rank + 1) * per, dtype=np.float64)
    if rank == 0:
        arr["b"][:] = 7
    arr = res.finish(arr)
    ok = np.array_equal(arr["a"], np.arange(1000, dtype=np.float64)) and (arr["b"] == 7).all()
    ok = ok and (arr["a"].flags.writeable == (rank == 0))
    import glob
    dist.barrier()
    ok = ok and not glob.glob(res.prefix + "*")        # unlinked: nothing left in /dev/shm
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shared_result_gloo_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shared_worker, args=(2, port, out), nprocs=2, join=True)
    assert all(out[r] for r in range(2))


def test_exchange_edges_gloo_world2():
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


def _pool_worker(rank, world, port, out):
    """Collective block choice of the shared result pool (graphtools_b200/hostpool.take_shared) under gloo: every rank
    maps the SAME segment, writes its slice, sees the others' slices; a segment is reused only when it is free on
    every rank, and a second one is created while any rank still holds a view of the first."""
    import gc
    from graphtools_b200 import hostpool as hp
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    n = 300_000
    lay, total = hp.layout([("a", n, np.float64), ("b", n, np.int32)])
    blk1 = hp.take_shared(total)
    ok &= blk1 is not None
    arr = blk1.carve(lay)
    half = n // world
    arr["a"][rank * half:(rank + 1) * half] = rank + 1.0
    arr["b"][rank * half:(rank + 1) * half] = rank + 7
    dist.barrier()
    for r in range(world):                                  # every rank sees every slice: one segment
        ok &= bool((arr["a"][r * half:(r + 1) * half] == r + 1.0).all())
        ok &= bool((arr["b"][r * half:(r + 1) * half] == r + 7).all())
    keep = arr["a"] if rank == 1 else None                  # rank 1 keeps a view: the segment is busy everywhere
    del arr
    gc.collect()
    ok &= (blk1.free == (rank != 1))
    blk2 = hp.take_shared(total)
    ok &= blk2 is not None and blk2 is not blk1
    a2 = blk2.carve(lay)
    del keep, a2
    gc.collect()
    dist.barrier()
    blk3 = hp.take_shared(total)                            # both free now: the first one is handed out again
    ok &= blk3 is blk1
    ok &= len(hp._shared) == 2
    ok &= not any(os.path.exists(p) for p in ["/dev/shm/gtbpool%d_%s_%d" % (os.getppid(), str(port), k) for k in (1, 2)])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shared_result_pool_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_pool_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
