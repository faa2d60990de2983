"""Page-locked host memory for the result matrices, recycled between builds.

``G.kernel`` / ``G.diff_op`` are host arrays (the graphtools contract: scipy CSR / ndarray on the caller's process).
Handing out fresh pageable memory costs a first-touch page fault per 4 KB and a staging copy per byte -- at 1M samples
that is 280 MB and ~30 ms on one GPU, and it does not shrink when the device work is sharded over eight.  Here the
arrays are carved out of page-locked blocks that the device writes by DMA directly:

* one process: blocks are ``cudaHostAlloc`` memory (torch pinned tensors);
* one process per GPU: a block is ONE shared-memory segment (``/dev/shm``) that every rank maps and registers with
  ``cudaHostRegister``; each rank copies its own row shard into its slice over its own PCIe link and all ranks view
  the complete arrays zero-copy.

A block returns to the pool when the last numpy view into it has been garbage-collected (``weakref.finalize`` on the
base array of every hand-out), so a caller that keeps ``G.kernel`` keeps its memory, and a loop that builds graph
after graph runs in the same pages.  With several ranks the choice of block is collective (all-reduce of the ranks'
"free" flags), so every rank maps the same segment.  ``GTB_HOST_POOL=0`` disables the pool (fresh pageable arrays,
staged copies); ``GTB_HOST_POOL_MB`` caps the page-locked bytes a process may hold (default 4096).
"""
import mmap
import os
import weakref

import numpy as np
import torch

_ALIGN = 256


def enabled():
    return os.environ.get("GTB_HOST_POOL", "1") != "0"


def _cap_bytes():
    return int(os.environ.get("GTB_HOST_POOL_MB", "4096")) << 20


def layout(specs):
    """specs = [(name, n_elements, numpy dtype)] -> ({name: (offset, count, dtype)}, total bytes), 256-byte aligned."""
    out, off = {}, 0
    for name, count, dt in specs:
        out[name] = (off, int(count), np.dtype(dt))
        off += (int(count) * np.dtype(dt).itemsize + _ALIGN - 1) // _ALIGN * _ALIGN
    return out, max(off, _ALIGN)


class _Block:
    """One page-locked block.  ``tensor`` / ``mm`` owns the memory (pinned torch tensor, or mmap of a shared segment)."""

    def __init__(self, nbytes, tensor=None, mm=None, registered=False):
        self.nbytes = nbytes
        self.tensor, self.mm, self.registered = tensor, mm, registered
        self.free = True

    def carve(self, lay, writeable=True):
        """{name: typed ndarray} according to ``lay``; the block is busy until every one of them (and every view
        derived from them) has been garbage-collected.  Each array is its own ``np.frombuffer`` over the block, NOT a
        view of one big uint8 array: scipy's constructors copy any array that is a view of a much larger ndarray
        (``_prune_array``), which would silently move the result out of the pool."""
        buf = memoryview(self.tensor.numpy()) if self.tensor is not None else self.mm
        holder = {"live": len(lay)}
        self.free = False
        out = {}
        for name, (off, count, dt) in lay.items():
            a = np.frombuffer(buf, dtype=dt, count=count, offset=off)
            weakref.finalize(a, _release, weakref.ref(self), holder)
            out[name] = a
        if not writeable:
            for a in out.values():
                a.flags.writeable = False
        return out

    def close(self):
        if self.mm is not None and self.registered:
            try:
                addr = np.frombuffer(self.mm, dtype=np.uint8, count=1).ctypes.data
                torch.cuda.cudart().cudaHostUnregister(addr)
            except Exception:
                pass
        self.tensor = self.mm = None


def _release(ref, holder):
    holder["live"] -= 1
    blk = ref()
    if blk is not None and holder["live"] == 0:
        blk.free = True


# ----------------------------------------------------------------------------------- one process
_local = []


def take_local(nbytes):
    """A free page-locked block of at least ``nbytes`` (None when the pool is off or full); ``block.carve(layout)``
    hands out the arrays."""
    if not enabled():
        return None
    for blk in _local:
        if blk.free and nbytes <= blk.nbytes <= 2 * nbytes + (1 << 20):
            return blk
    held = sum(b.nbytes for b in _local)
    size = (nbytes + nbytes // 8 + (1 << 20) - 1) >> 20 << 20          # headroom: the next build's nnz differs a little
    # drop free blocks that no longer fit the requests being made
    for blk in [b for b in _local if b.free]:
        if held + size > _cap_bytes():
            _local.remove(blk)
            held -= blk.nbytes
            blk.close()
    if held + size > _cap_bytes():
        return None
    try:
        t = torch.empty((size,), dtype=torch.uint8, pin_memory=True)
    except RuntimeError:
        return None
    blk = _Block(size, tensor=t)
    _local.append(blk)
    return blk


def d2h_async(dev_tensor, host_array):
    """Device tensor -> page-locked host array of the same byte size, asynchronous on the current stream (DMA straight
    into the destination; the caller synchronises the stream before reading)."""
    nbytes = dev_tensor.numel() * dev_tensor.element_size()
    assert host_array.nbytes == nbytes, (host_array.nbytes, nbytes)
    if nbytes == 0:
        return
    dst = torch.from_numpy(host_array.reshape(-1).view(np.uint8))
    dst.copy_(dev_tensor.contiguous().view(-1).view(torch.uint8), non_blocking=True)


# ----------------------------------------------------------------------------------- one process per GPU
_shared = []
_seq = [0]


def take_shared(nbytes, group=None):
    """Collective.  A page-locked shared segment of at least ``nbytes`` that every rank of the group maps (None on
    every rank when the pool is off, /dev/shm is too small or registration fails anywhere)."""
    import torch.distributed as dist
    if not enabled():
        return None
    rank = dist.get_rank(group)
    # (a gloo group on CPU -- the host-logic tests -- agrees through CPU tensors and skips the page-locking)
    on_gpu = torch.cuda.is_available() and "nccl" in str(dist.get_backend(group))
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    # blocks are created collectively, so the lists agree across ranks; a block is usable when it is free EVERYWHERE
    if _shared:
        flags = torch.tensor([int(b.free and nbytes <= b.nbytes <= 2 * nbytes + (1 << 20)) for b in _shared],
                             dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)
        flags = flags.tolist()
        for blk, ok in zip(_shared, flags):
            if ok:
                return blk
        # retire blocks that are free everywhere but have the wrong size
        free_all = torch.tensor([int(b.free) for b in _shared], dtype=torch.int32, device=dev)
        dist.all_reduce(free_all, op=dist.ReduceOp.MIN, group=group)
        for blk, f in list(zip(_shared, free_all.tolist())):
            if f:
                _shared.remove(blk)
                blk.close()
    size = (nbytes + nbytes // 8 + (1 << 20) - 1) >> 20 << 20
    ok_local = 1
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize < size + (64 << 20) or sum(b.nbytes for b in _shared) + size > _cap_bytes():
            ok_local = 0
    except OSError:
        ok_local = 0
    ok = torch.tensor([ok_local], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if not int(ok.item()):
        return None
    _seq[0] += 1
    path = "/dev/shm/gtbpool%d_%s_%d" % (os.getppid(), os.environ.get("MASTER_PORT", "0"), _seq[0])
    if rank == 0:
        if os.path.exists(path):
            os.unlink(path)
        fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
        os.ftruncate(fd, size)
        os.close(fd)
    dist.barrier(group=group)
    fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, size)
    os.close(fd)
    # every rank touches (allocates) its own 1/world of the pages, in parallel, then registers the whole mapping
    world = dist.get_world_size(group)
    arr = np.frombuffer(mm, dtype=np.uint8, count=size)
    per = (size + world - 1) // world
    arr[rank * per: min(size, (rank + 1) * per)] = 0
    dist.barrier(group=group)
    registered = 1
    if on_gpu:
        try:
            rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, size, 0)
            if int(rc) != 0:
                registered = 0
        except Exception:
            registered = 0
    del arr
    reg = torch.tensor([registered], dtype=torch.int32, device=dev)
    dist.all_reduce(reg, op=dist.ReduceOp.MIN, group=group)
    if rank == 0:
        os.unlink(path)                        # the mappings keep the segment alive; nothing is left behind
    blk = _Block(size, mm=mm, registered=bool(registered) and on_gpu)
    if not int(reg.item()):
        blk.close()
        return None
    _shared.append(blk)
    return blk
