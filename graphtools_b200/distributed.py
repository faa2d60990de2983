"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

Partitioning: query rows are block-partitioned over the ranks, the reference set is
replicated (1M x 100 float32 = 400 MB << 180 GB).  Distance/top-k, float64 refine and CSR emission are
independent per row, so each rank builds the raw kernel rows of its shard with no communication.
The one exchange step is ahead of symmetrisation: the raw CSR row shards are all-gathered (variable
length -> padded to the longest shard), after which symmetrise / normalise are local.  Because every row
sees the whole reference set, the result is bit-identical for any number of ranks.
"""
import torch
import torch.distributed as dist


def active():
    """True when the graph build should shard its query rows over the ranks of the default process
    group (one process per GPU launched by torchrun; set GTB_DISTRIBUTED=0 to opt out)."""
    import os
    return (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            and os.environ.get("GTB_DISTRIBUTED", "1") != "0")


MIN_ROWS_PER_RANK = 128   # out-of-sample query sets smaller than this per rank are not worth sharding


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def shard_bounds(n, world, rank):
    """Contiguous row range [lo, hi) owned by ``rank``; multiples of 128 rows so query tiles stay full."""
    tiles = (n + 127) // 128
    per = (tiles + world - 1) // world
    lo = min(n, rank * per * 128)
    hi = min(n, (rank + 1) * per * 128)
    return lo, hi


def _allgather_padded(t, length, group=None):
    """All-gather 1-D tensors of different lengths (``length`` = list of per-rank lengths)."""
    world = dist.get_world_size(group)
    m = max(max(length), 1)
    buf = torch.zeros((m,), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, length)])


def allgather_csr_rows(row_len, indices, data, n_rows_per_rank, scan_fn, group=None):
    """Row-sharded CSR pieces -> the full CSR on every rank.

    row_len [m_r] int32 (entries per local row), indices / data [nnz_r]; ``n_rows_per_rank`` = list of
    shard heights; ``scan_fn(int32 tensor) -> int64 tensor [n+1]`` builds the row pointers (the CUDA
    scan in production).  Returns (indptr [N+1] int64, indices [nnz], data [nnz])."""
    world = dist.get_world_size(group)
    nnz_local = torch.tensor([indices.shape[0]], dtype=torch.int64, device=indices.device)
    nnz_all = [torch.empty_like(nnz_local) for _ in range(world)]
    dist.all_gather(nnz_all, nnz_local, group=group)
    nnz_all = [int(x.item()) for x in nnz_all]
    full_len = _allgather_padded(row_len, list(n_rows_per_rank), group)
    full_idx = _allgather_padded(indices, nnz_all, group)
    full_val = _allgather_padded(data, nnz_all, group)
    return scan_fn(full_len.contiguous()), full_idx, full_val


def owner_of(cols, bounds):
    """Rank that owns each column/row index under ``shard_bounds`` (bounds = [(lo, hi)] per rank)."""
    world = len(bounds)
    per = max(bounds[0][1] - bounds[0][0], 1)
    return torch.clamp(cols.to(torch.int64) // per, max=world - 1)


def route_edges_to_column_owner(row_len, indices, data, lo, bounds, group=None):
    """The exchange step of the sharded symmetrisation: every raw edge (i, j, w) of this rank's rows is sent
    to owner(j) with one NCCL all-to-all per field (variable splits).  Returns the edges received by this rank
    as the TRANSPOSED entries of its own rows -- a CSR over the local rows (row_len_t [m] int32, cols [k] int32
    = source row ids i, vals [k] float64), column-sorted -- ready to be merged with the local raw rows.

    Pure plumbing on torch tensors (works on CPU tensors with gloo for the tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = indices.device
    m = row_len.shape[0]
    n_total = bounds[-1][1]
    rows = torch.repeat_interleave(torch.arange(lo, lo + m, device=dev, dtype=torch.int64), row_len.to(torch.int64))
    dest = owner_of(indices, bounds)
    order = torch.argsort(dest, stable=True)
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    in_split = send_counts.tolist()
    out_split = recv_counts.tolist()
    k = int(sum(out_split))

    def a2a(t):
        out = torch.empty((k,), dtype=t.dtype, device=dev)
        dist.all_to_all_single(out, t[order].contiguous(), output_split_sizes=out_split, input_split_sizes=in_split,
                               group=group)
        return out

    src_row = a2a(rows.to(torch.int32))          # i  (becomes the column of the transposed entry)
    dst_col = a2a(indices.to(torch.int32))       # j  (a row of this rank)
    val = a2a(data)
    my_lo = bounds[rank][0]
    my_m = bounds[rank][1] - my_lo
    local_row = dst_col.to(torch.int64) - my_lo
    key = local_row * n_total + src_row.to(torch.int64)
    perm = torch.argsort(key)
    row_len_t = torch.bincount(local_row, minlength=my_m).to(torch.int32)
    return row_len_t, src_row[perm].contiguous(), val[perm].contiguous()
