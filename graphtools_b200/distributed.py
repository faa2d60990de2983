"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

Partitioning: query rows are block-partitioned over the ranks in multiples of 128.

1. **Reference set**: rank r uploads only ITS rows of X from the host and the shards are assembled with one
   NCCL all-gather (``upload_sharded``) -- the host-to-device traffic of a build is |X| in total, not |X| per rank.
2. Distance / top-k, float64 refine and CSR emission are independent per row: no communication.
3. **Edge exchange** ahead of symmetrisation: every raw edge (i, j, w) goes to the rank that owns column j.  The
   edges are bucketed by two kernels (csrc/symm.cu ``route_count`` / ``route_fill``: rows are column-sorted, so the
   entries bound for one rank are contiguous in each row -- a count, one scan and one scatter, no sort), sent as packed
   16-byte records with ONE all-to-all, and turned into the rows of the transposed matrix on the receiver by the
   histogram / scan / scatter / per-row-sort kernels of the single-GPU transpose.  One host synchronisation (the
   split sizes) per exchange.
4. Symmetrise / normalise run per shard with the same merge kernel as the single-GPU build, so K and P are
   bit-identical for any number of ranks.  The result stays ROW-SHARDED in HBM; the full matrix is assembled
   only when somebody asks for it: on the device by ``allgather_csr_rows`` (landmark / MNN consumers), on the host
   through a shared-memory segment that every rank fills with its own shard in parallel (``SharedResult``).
"""
import mmap
import os

import numpy as np
import torch
import torch.distributed as dist


def active():
    """True when the graph build should shard its query rows over the ranks of the default process
    group (one process per GPU launched by torchrun; set GTB_DISTRIBUTED=0 to opt out)."""
    return (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            and os.environ.get("GTB_DISTRIBUTED", "1") != "0")


MIN_ROWS_PER_RANK = 128   # out-of-sample query sets smaller than this per rank are not worth sharding


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def shard_bounds(n, world, rank):
    """Contiguous row range [lo, hi) owned by ``rank``; multiples of 128 rows so query tiles stay full."""
    per = rows_per_rank(n, world)
    lo = min(n, rank * per)
    hi = min(n, (rank + 1) * per)
    return lo, hi


def rows_per_rank(n, world):
    tiles = (n + 127) // 128
    return max(1, (tiles + world - 1) // world) * 128


def _allgather_padded(t, length, group=None):
    """All-gather 1-D tensors of different lengths (``length`` = list of per-rank lengths)."""
    world = dist.get_world_size(group)
    m = max(max(length), 1)
    buf = torch.zeros((m,), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, length)])


def allgather_counts(value, device, group=None):
    """One integer per rank -> list of ints on every rank."""
    world = dist.get_world_size(group)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.empty((world,), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(x) for x in out.tolist()]


def allgather_csr_rows(row_len, indices, data, n_rows_per_rank, scan_fn, group=None):
    """Row-sharded CSR pieces -> the full CSR on every rank.

    row_len [m_r] int32 (entries per local row), indices / data [nnz_r]; ``n_rows_per_rank`` = list of
    shard heights; ``scan_fn(int32 tensor) -> int64 tensor [n+1]`` builds the row pointers (the CUDA
    scan in production).  Returns (indptr [N+1] int64, indices [nnz], data [nnz])."""
    nnz_all = allgather_counts(indices.shape[0], indices.device, group)
    full_len = _allgather_padded(row_len, list(n_rows_per_rank), group)
    full_idx = _allgather_padded(indices, nnz_all, group)
    full_val = _allgather_padded(data, nnz_all, group)
    return scan_fn(full_len.contiguous()), full_idx, full_val


def owner_of(cols, bounds):
    """Rank that owns each column/row index under ``shard_bounds`` (bounds = [(lo, hi)] per rank)."""
    world = len(bounds)
    per = max(bounds[0][1] - bounds[0][0], 1)
    return torch.clamp(cols.to(torch.int64) // per, max=world - 1)


# --------------------------------------------------------------------------- reference-set assembly
def upload_sharded(X_host, dtype, device, group=None):
    """Host rows -> full device tensor [n, d]: this rank copies rows [lo, hi) only (pinned sources go at full
    PCIe rate), NCCL all-gathers the shards (SURVEY 8e collective 1).  ``X_host``: numpy array or CPU tensor, the
    same on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Xh = X_host if isinstance(X_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X_host))
    n, d = Xh.shape
    per = rows_per_rank(n, world)
    lo, hi = shard_bounds(n, world, rank)
    full = torch.empty((world * per, d), dtype=dtype, device=device)
    mine = full[rank * per: rank * per + (hi - lo)]
    if hi > lo:
        src = Xh[lo:hi]
        if src.dtype != dtype:
            src = src.to(dtype)
        mine.copy_(src, non_blocking=True)
    dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per].clone(), group=group)
    return full[:n]


# --------------------------------------------------------------------------- edge exchange
def cuda_bucket_edges(indptr, indices, data, lo, per, world):
    """(send records [nnz, 2] int64 = packed {int32 i, int32 j, float64 w}, send_counts [world] int64 on the device):
    the raw edges of this rank's rows in (destination, row, column) order (csrc/symm.cu route_count / route_fill)."""
    from . import _engine as E
    from . import pipeline
    m = indptr.shape[0] - 1
    nnz = indices.shape[0]
    dev = indices.device
    send = torch.empty((nnz, 2), dtype=torch.int64, device=dev)
    if m == 0 or nnz == 0:
        return send, torch.zeros((world,), dtype=torch.int64, device=dev)
    cnt = torch.empty((world * m,), dtype=torch.int32, device=dev)
    E.call("gtb_route_count", indptr, indices, m, per, world, cnt)
    pos = pipeline.exclusive_scan(cnt)
    E.call("gtb_route_fill", indptr, indices, data, m, lo, per, world, cnt, pos, send)
    edges = pos[torch.arange(0, world + 1, device=dev) * m]
    return send, edges[1:] - edges[:-1]


def exchange_edges(indptr, indices, data, lo, bounds, bucket_fn=cuda_bucket_edges, group=None):
    """The exchange step of the sharded symmetrisation: every raw edge (i, j, w) of this rank's rows is sent to
    owner(j).  Returns the records received by this rank, [k, 2] int64 (packed {int32 i, int32 j, float64 w}): the
    transposed entries of its own rows, in arrival order.  One all-to-all for the split sizes, one for the records,
    one host synchronisation in between (the sizes of the receive buffer)."""
    world = dist.get_world_size(group)
    per = max(bounds[0][1] - bounds[0][0], 1)
    send, send_counts = bucket_fn(indptr, indices, data, lo, per, world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    splits = torch.stack([send_counts, recv_counts]).cpu()          # the one host sync
    in_split, out_split = splits[0].tolist(), splits[1].tolist()
    recv = torch.empty((int(sum(out_split)), 2), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    return recv


def unpack_records(rec):
    """[k, 2] int64 records -> (i int32 [k], j int32 [k], w float64 [k]) (host-side helper for tests / debugging)."""
    if rec.shape[0] == 0:
        z = torch.zeros((0,), dtype=torch.int32, device=rec.device)
        return z, z.clone(), torch.zeros((0,), dtype=torch.float64, device=rec.device)
    ij = rec[:, 0].contiguous().view(torch.int32).view(-1, 2)
    return ij[:, 0].contiguous(), ij[:, 1].contiguous(), rec[:, 1].contiguous().view(torch.float64)


# --------------------------------------------------------------------------- host materialisation
_seq = [0]


class SharedResult:
    """Host arrays of a row-sharded result assembled in ONE shared-memory segment per array: rank 0 creates
    /dev/shm/gtb<job>_<seq>_<name>, every rank maps it and copies its own shard into its slice (device->host in
    parallel over the ranks' own PCIe links; first-touch page faults in parallel too), after a barrier every rank
    holds a zero-copy numpy view of the complete array.  Rank 0's views are writable, the others read-only.  The
    segment is unlinked as soon as everybody has mapped it, so nothing is left behind; the mapping lives as long as
    the arrays do."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        _seq[0] += 1
        self.prefix = "/dev/shm/gtb%d_%s_%d" % (os.getppid(), os.environ.get("MASTER_PORT", "0"), _seq[0])
        self.maps = {}
        self.paths = []

    @staticmethod
    def available(nbytes):
        try:
            st = os.statvfs("/dev/shm")
            return st.f_bavail * st.f_frsize > nbytes + (64 << 20)
        except OSError:
            return False

    def create(self, specs):
        """specs = {name: (n_elements, numpy dtype)}; collective."""
        if self.rank == 0:
            for name, (count, dt) in specs.items():
                path = "%s_%s" % (self.prefix, name)
                if os.path.exists(path):
                    os.unlink(path)
                fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
                os.ftruncate(fd, max(1, count * np.dtype(dt).itemsize))
                os.close(fd)
        dist.barrier(group=self.group)
        out = {}
        for name, (count, dt) in specs.items():
            path = "%s_%s" % (self.prefix, name)
            fd = os.open(path, os.O_RDWR)
            mm = mmap.mmap(fd, max(1, count * np.dtype(dt).itemsize))
            os.close(fd)
            self.paths.append(path)
            out[name] = np.frombuffer(mm, dtype=dt, count=count)
        return out

    def finish(self, arrays):
        """Collective: everybody has written its slice -> unlink, hand out the views."""
        dist.barrier(group=self.group)
        if self.rank == 0:
            for path in self.paths:
                os.unlink(path)
        else:
            for a in arrays.values():
                a.flags.writeable = False
        return arrays
