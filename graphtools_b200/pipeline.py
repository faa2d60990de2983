"""Device pipeline: the order in which the C-ABI kernels are chained for one graph build.

search operand -> fused distance/top-S -> float64 refine + certification -> (radius pass for
uncertified rows) -> CSR emission -> symmetrise -> row-normalise.

Everything stays in HBM as torch tensors; the two host synchronisations per kernel build are the
reads of `number of uncertified rows` and `nnz` (sizes of the next allocations).
"""
import math

import numpy as np
import torch

from . import _engine as E

SYM_MODES = {"+": 0, "*": 1, "mnn": 2, None: 3}
METRIC_ALIASES = {"manhattan": "cityblock", "l1": "cityblock", "l2": "euclidean"}
METRIC_CODE = {"euclidean": 0, "cosine": 1, "cityblock": 2}
AUTO_TC = "tch1"         # tensor-core flavour picked by impl="auto": "tch1" (fp16, one product, seeded), "tch" (fp16x2), "tc16" (bf16x3) or "tc" (3xTF32)
BALL_CAP = 8192          # longest radius-pass row handled by refine_ball (shared-memory sort)
_STATS = {}


class RowTooLong(NotImplementedError):
    """A row has more neighbours inside the kernel radius than the radius-pass refine handles (BALL_CAP)."""


def stats():
    """Counters of the last kernel build (rows sent to the radius pass, passes, nnz)."""
    return dict(_STATS)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=_dev())


def _zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=_dev())


def to_device_f32(X):
    """Host array / tensor -> contiguous float32 CUDA tensor."""
    E.require_cuda()
    if isinstance(X, torch.Tensor):
        return X.to(device=_dev(), dtype=torch.float32).contiguous()
    X = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
    t = torch.from_numpy(X)
    return t.to(_dev(), non_blocking=False)


def to_device(X):
    """Host array / tensor -> contiguous CUDA tensor that keeps float64 inputs in float64 (the exact
    distances are then evaluated on the original float64 rows, e.g. PCA output); everything else becomes
    float32."""
    E.require_cuda()
    if isinstance(X, torch.Tensor):
        dt = torch.float64 if X.dtype == torch.float64 else torch.float32
        return X.to(device=_dev(), dtype=dt).contiguous()
    X = np.asarray(X)
    dt = np.float64 if X.dtype == np.float64 else np.float32
    return torch.from_numpy(np.ascontiguousarray(X, dtype=dt)).to(_dev())


_D2H_CHUNK = 32 << 20      # bytes per staging buffer
_D2H_RING = 3
_d2h_state = {}


def _d2h_ring():
    """Page-locked staging ring, a side stream and a small thread pool, created once per process: the result
    arrays (K / P values, indices: ~350 MB at 1M samples) are handed to the caller, so they must live in ordinary
    pageable memory, but a direct pageable cudaMemcpy runs at ~2.6 GB/s (first-touch page faults serialised with
    the copy)."""
    if not _d2h_state:
        from concurrent.futures import ThreadPoolExecutor
        _d2h_state["bufs"] = [torch.empty((_D2H_CHUNK,), dtype=torch.uint8, pin_memory=True) for _ in range(_D2H_RING)]
        _d2h_state["stream"] = torch.cuda.Stream()
        _d2h_state["pool"] = ThreadPoolExecutor(max_workers=4)
    return _d2h_state["bufs"], _d2h_state["stream"], _d2h_state["pool"]


def d2h_into(t, dst_array):
    """Device tensor -> an existing host numpy array of the same size and dtype (e.g. a slice of a shared-memory
    segment), through the pinned staging ring."""
    assert dst_array.nbytes == t.numel() * t.element_size(), (dst_array.nbytes, t.numel(), t.element_size())
    if t.numel() == 0:
        return
    d2h_pinned(t, dst_array.reshape(-1).view(np.uint8))


def d2h_pinned(t, dst=None):
    """Device->host copy into a fresh pageable tensor (or into ``dst``, a uint8 numpy view of the destination).
    Large tensors go through the pinned staging ring in 32 MB chunks: the DMA of chunk i+1 overlaps the host-side
    copy of chunk i, which four threads perform in parallel (so the first-touch page faults of the destination are
    taken in parallel too)."""
    nbytes = t.numel() * t.element_size()
    if nbytes < (8 << 20) or not t.is_cuda:
        if dst is None:
            return t.cpu()
        dst[:] = t.contiguous().view(-1).view(torch.uint8).cpu().numpy()
        return None
    bufs, side, pool = _d2h_ring()
    src = t.contiguous().view(-1).view(torch.uint8)
    out = None
    if dst is None:
        out = torch.empty(t.shape, dtype=t.dtype)
        dst = out.view(-1).view(torch.uint8).numpy()
    side.wait_stream(torch.cuda.current_stream())
    chunks = [(off, min(_D2H_CHUNK, nbytes - off)) for off in range(0, nbytes, _D2H_CHUNK)]
    events = [None] * len(chunks)
    pending = [[] for _ in range(_D2H_RING)]           # host copies still reading staging buffer r

    def issue(i):
        off, n = chunks[i]
        r = i % _D2H_RING
        for f in pending[r]:
            f.result()
        pending[r] = []
        with torch.cuda.stream(side):
            bufs[r][:n].copy_(src[off:off + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        events[i] = ev

    def drain(i):
        off, n = chunks[i]
        r = i % _D2H_RING
        events[i].synchronize()
        hb = bufs[r].numpy()
        step = (n + 3) // 4
        for a in range(0, n, step):
            b = min(n, a + step)
            pending[r].append(pool.submit(np.copyto, dst[off + a:off + b], hb[a:b]))

    for i in range(min(_D2H_RING - 1, len(chunks))):
        issue(i)
    for i in range(len(chunks)):
        if i + _D2H_RING - 1 < len(chunks):
            issue(i + _D2H_RING - 1)
        drain(i)
    for r in range(_D2H_RING):
        for f in pending[r]:
            f.result()
    t.record_stream(side)
    return out


class SearchOperand:
    """k-major centred copy of a point set + squared norms (csrc/prep.cu)."""

    def __init__(self, X, mean=None, metric="euclidean"):
        """X: the ORIGINAL rows on the device, float32 or float64 (kept for the exact distances).  The search
        copy is always float32: float64 inputs are centred in float64 first and rounded once, so the
        rounding error of the fast pass stays relative to the centred magnitudes (same bound E).

        metric="cosine": the search copy is built from the row-normalised data (normalised and centred in float64,
        rounded once), on which squared Euclidean distance = 2 x cosine distance; the exact stage evaluates
        1 - x.y / (|x||y|) on the original rows."""
        n, d = X.shape
        self.X, self.n, self.d = X, n, d
        self.metric = metric
        self.is64 = X.dtype == torch.float64
        self.n_pad = (n + 127) // 128 * 128
        self.d_pad = (d + 7) // 8 * 8
        metric = METRIC_ALIASES.get(metric, metric)
        self.metric = metric
        if metric not in ("euclidean", "cosine", "cityblock"):
            raise NotImplementedError("metric {!r} is not supported by the CUDA search".format(metric))
        if metric == "cityblock":
            # |x - y| is translation invariant and has no GEMM form: the search copy is the float32 data itself,
            # NOT centred, so that the float32 accumulation of sum |x - y| stays accurate relative to the distance
            # (float64 inputs are rounded once; that rounding enters the certification bound through the L1 norms)
            self.Xs = X.to(torch.float32).contiguous() if self.is64 else X
            self._kmean = None
            mean = torch.zeros((d,), dtype=torch.float32, device=X.device) if mean is None else mean
        elif self.is64 or metric == "cosine":
            rows = X.to(torch.float64)
            if metric == "cosine":
                rows = rows / rows.norm(dim=1, keepdim=True)
            if mean is None:
                mean = rows.mean(dim=0)
            self.Xs = (rows - mean.to(torch.float64)).to(torch.float32).contiguous()
            self._kmean = None                       # already centred
            del rows
        else:
            if mean is None:
                ws = _empty((E.lib().gtb_col_mean_ws_doubles(d),), torch.float64)
                mean = _empty((d,), torch.float32)
                E.call("gtb_col_mean", X, n, d, ws, mean)
            self.Xs = X
            self._kmean = mean.to(torch.float32)
        self.mean = mean
        self._simt = None
        self._tc = {}
        self._n2_tc = None
        self._maxnorm = None
        self._maxnorm_host = None

    # -- CUDA-core operand: k-major [d_pad][n_pad] + norms
    def simt(self):
        if self._simt is None:
            XT = _empty((self.d_pad, self.n_pad), torch.float32)
            n2 = _empty((self.n_pad,), torch.float32)
            mx = _empty((1,), torch.float32)
            E.call("gtb_prepare_operand", self.Xs, self.n, self.d, self._kmean, XT, self.n_pad, self.d_pad, n2, mx)
            self._simt = (XT, n2)
            if self._maxnorm is None:
                self._maxnorm = mx
        return self._simt

    @property
    def XT(self):
        return self.simt()[0]

    @property
    def n2(self):
        return self.simt()[1]

    # -- tensor-core operand: row-major [n_pad][Kp] hi/lo pair for one role (0 query, 1 reference);
    #    dtype 0 = tf32 pairs stored as float32 (3xTF32), dtype 1 = bfloat16 pairs (bf16x3)
    def kp(self, dtype=0):
        # rows are whole 32-byte k-steps (16 bf16 / 8 tf32 elements): SWIZZLE_128B blocks plus SWIZZLE_32B
        # tail blocks.  (Padding bf16 rows to whole 128-byte blocks was measured slower: 699 vs 682 ms at
        # d = 100 -- the extra MMAs cost more than the tail blocks.)
        import os
        step = 16 if dtype else 8
        step = int(os.environ.get("GTB_KP_STEP", step))      # experiments: 64 (bf16) / 32 (tf32) = whole 128-byte blocks
        # float16 operands keep one more column: the low part of |y|^2 (shared by the two- and one-product flavours)
        return (self.d + 1 + (1 if dtype >= 2 else 0) + step - 1) // step * step

    @property
    def Kp(self):
        return self.kp(0)

    def tc_ok(self, dtype=0):
        if self.metric == "cityblock":
            return False                                  # no GEMM form: CUDA-core L1 kernel only
        # resident query tile: 13 tf32 / 8 bf16 k-steps; the fp16x2 flavour streams longer rows in chunks (d <= 510)
        return self.kp(dtype) // (16 if dtype else 8) <= {0: 13, 1: 8, 2: 32, 3: 8}[dtype]

    def l1_norms(self):
        """(per-row L1 norms float32 [n], max) of the search copy -- bound on its rounding for float64 inputs."""
        if getattr(self, "_l1", None) is None:
            n1 = self.Xs.abs().sum(dim=1, dtype=torch.float64)
            self._l1 = (n1.to(torch.float32).contiguous(), float(n1.max().item()))
        return self._l1

    def tc(self, role, dtype=0, scale=1.0):
        """(hi, lo, norm2) of one role; dtype 0 = tf32 pairs in float32, 1 = bfloat16 pairs, 2 = float16 pairs of the
        data scaled by ``scale`` (norm2 stays unscaled)."""
        dtype = min(dtype, 2)                             # the one-product flavour reads the fp16x2 hi arrays
        key = (role, dtype, float(scale))
        if key not in self._tc:
            Kp = self.kp(dtype)
            st = (torch.float32, torch.bfloat16, torch.float16)[dtype]
            hi = _empty((self.n_pad, Kp), st)
            lo = _empty((self.n_pad, Kp), st)
            if getattr(self, "_n2_tc", None) is not None:
                # the norms are on hand (norm_max, or the other role): split only
                n2 = self._n2_tc
                E.call("gtb_split_operand_tc", self.Xs, self.n, self.d, self._kmean, role, hi, lo, self.n_pad, Kp,
                       dtype, float(scale), n2)
            else:
                n2 = _empty((self.n_pad,), torch.float32)
                mx = _empty((1,), torch.float32)
                E.call("gtb_prepare_operand_tc", self.Xs, self.n, self.d, self._kmean, role, hi, lo, self.n_pad, Kp,
                       dtype, float(scale), n2, mx)
                self._n2_tc = n2
                if self._maxnorm is None:
                    self._maxnorm = mx
            self._tc[key] = (hi, lo, n2)
        return self._tc[key]

    def norm_max(self):
        """max squared norm of the centred rows WITHOUT building an operand (the fp16 flavour needs it to choose its
        scale before the split); one pass over X and one host read."""
        if self._maxnorm_host is None and self._maxnorm is None:
            n2 = _empty((self.n_pad,), torch.float32)
            mx = _empty((1,), torch.float32)
            E.call("gtb_row_norms", self.Xs, self.n, self.d, self._kmean, self.n_pad, n2, mx)
            self._maxnorm = mx
            self._n2_tc = n2
        return self.maxnorm

    @property
    def maxnorm(self):
        if self._maxnorm_host is None:
            if self._maxnorm is None:
                self.simt()
            self._maxnorm_host = float(self._maxnorm.item())
        return self._maxnorm_host


_NP_OF = {torch.int32: np.int32, torch.int64: np.int64, torch.float64: np.float64, torch.float32: np.float32}


def _d2h_pooled(named):
    """[(name, device tensor)] -> {name: numpy array} carved out of ONE recycled page-locked block (hostpool.py) and
    filled by DMA -- no staging copy, no first-touch page faults after the first build.  None when the pool is off,
    full, or the arrays are small (the plain path is as good)."""
    from . import hostpool
    if not all(t.is_cuda for _, t in named) or sum(t.numel() * t.element_size() for _, t in named) < (8 << 20):
        return None
    lay, total = hostpool.layout([(nm, t.numel(), _NP_OF[t.dtype]) for nm, t in named])
    blk = hostpool.take_local(total)
    if blk is None:
        return None
    out = blk.carve(lay)
    for nm, t in named:
        hostpool.d2h_async(t, out[nm])
    torch.cuda.current_stream().synchronize()
    return out


class DeviceCSR:
    """CSR in HBM: int64 indptr [n+1], int32 indices, float64 data."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, tuple(shape)

    @property
    def nnz(self):
        return int(self.indices.shape[0])

    def _host_structure(self):
        """(indices, indptr int32) as numpy arrays, copied device->host once through pinned buffers and
        shared by every scipy view of this matrix (K, P, transitions...)."""
        if getattr(self, "_host_struct", None) is None:
            n1 = self.indptr.shape[0]
            ip32 = _empty((n1,), torch.int32)
            E.call("gtb_cast_indptr", self.indptr, n1, ip32)
            got = _d2h_pooled([("indices", self.indices), ("indptr", ip32)])
            if got is not None:
                self._host_struct = (got["indices"], got["indptr"])
            else:
                idx_h, ip_h = d2h_pinned(self.indices), d2h_pinned(ip32)
                torch.cuda.current_stream().synchronize()
                self._host_struct = (idx_h.numpy(), ip_h.numpy())
        return self._host_struct

    def to_scipy(self, data=None):
        from scipy import sparse
        src = self.data if data is None else data
        got = _d2h_pooled([("vals", src)])
        vals = got["vals"] if got is not None else d2h_pinned(src).numpy()
        indices, indptr = self._host_structure()
        torch.cuda.current_stream().synchronize()
        M = sparse.csr_matrix((vals, indices, indptr), shape=self.shape)
        M.has_sorted_indices = True
        M.has_canonical_format = True
        return M


def exclusive_scan(counts):
    n = counts.shape[0]
    out = _empty((n + 1,), torch.int64)
    ws = _empty((E.lib().gtb_scan_ws_elems(n),), torch.int64)
    E.call("gtb_exclusive_scan", counts, n, out, ws)
    return out


def choose_S(knn_eff, binary):
    """Candidate-list length of the fused top-k: the reference's first search is
    knn * search_multiplier = 36 wide (graphs.py:882); we keep a little more so that the
    certification margin rarely sends a row to the radius pass."""
    need = knn_eff + (4 if binary else 8)
    for S in ((16, 32, 48, 64, 128) if binary else (48, 64, 128)):
        if need <= S:
            return S
    raise NotImplementedError(
        "knn={} needs a candidate list longer than 128; not supported by the fused top-k yet".format(knn_eff))


def seed_stride():
    """GTB_TC_SEED_STRIDE: the seed sweep of the one-product flavour visits every stride-th reference tile (1 = off)."""
    import os
    return int(os.environ.get("GTB_TC_SEED_STRIDE", "16"))


def seedable(ref):
    """A strided sample needs enough tiles to carry the order statistic: 32 sampled tiles at least."""
    st = seed_stride()
    return st > 1 and ref.n_pad // 128 >= 32 * st


def tc_cluster():
    """GTB_TC_CLUSTER = 1 | 2 | 4: CTAs per cluster sharing each reference tile by TMA multicast."""
    import os
    return int(os.environ.get("GTB_TC_CLUSTER", "2"))


def default_impl():
    """GTB_SEARCH_IMPL = tc (3xTF32 tensor cores) | tc16 (bf16x3 tensor cores) | tch (fp16x2 tensor cores: two
    products, wider certified bound) | tch1 (fp16, ONE product, seeded thresholds, lists of 64) | simt (fp32 CUDA
    cores) | auto (default: tensor cores whenever the operand fits)."""
    import os
    return os.environ.get("GTB_SEARCH_IMPL", "auto")


def eps_rel_tc(d):
    """Same bound for the 3xTF32 tensor-core pass: dropped lo*lo terms and the tf32 rounding of the lo
    parts contribute 3 * 2^-22, float32 accumulation in the tensor core at most Kp * 2^-23 (truncation);
    doubled."""
    return 4.0 * (d + 16) * 2.0 ** -24


def eps_rel_tc16(d):
    """bf16x3 (hi*hi + hi*lo + lo*hi, 16 mantissa bits kept): dropped terms <= (2^-16 + 2^-18) |a||b| per
    product, sum |a||b| <= 2 (|x~|^2 + max|y~|^2); float32 accumulation as for tf32."""
    return 2.0 * (2.0 ** -16 + 2.0 ** -18) + 2.0 * (d + 32) * 2.0 ** -24


def eps_rel_l1(d):
    """Cityblock pass: float32 subtraction and accumulation of d terms |x_k - y_k| -- relative error of the
    DISTANCE at most (d + 1) * 2^-24 (all terms non-negative, no cancellation), doubled."""
    return 2.0 * (d + 2) * 2.0 ** -24


def eps_rel_tch(d):
    """fp16x2 (A_hi B_hi + A_hi B_lo on float16 pairs): the dropped query low part is <= 2^-11 |a| per element, so
    the dropped product is <= 2^-11 sum |a||b| <= 2^-11 |x~| |2 y~| <= 2^-11 (|x~|^2 + |y~|^2); the reference operand
    keeps 22 bits (2^-22, twice for the two parts), float32 accumulation as for the other flavours, 1e-6 covers fp16
    subnormals of the low parts at the chosen scale."""
    return 2.0 ** -11 + 2.0 ** -20 + 2.0 * (d + 32) * 2.0 ** -24 + 1e-6


def fp16_scale(maxnorm):
    """Power of two s with s^2 * maxnorm <= gtb_tc_fp16_maxnorm(): scaled data fits float16 with headroom."""
    limit = float(E.lib().gtb_tc_fp16_maxnorm())
    if not (maxnorm > 0):
        return 1.0
    return 2.0 ** math.floor(0.5 * math.log2(limit / maxnorm))


TC_DTYPE = {"tc": 0, "tc16": 1, "tch": 2, "tch1": 3}


def eps_rel_tch1(d):
    """fp16x1 (A_hi B_hi only, float16): both operands are rounded to 11 bits, |delta(2 x.y)| <= 2 (2 u + u^2)
    sum |x_k||y_k| <= 2^-10 (1 + 2^-12) (|x~|^2 + |y~|^2) with u = 2^-11; |y|^2 keeps 22 bits (hi + spare-column low
    part); float32 accumulation and fp16 subnormals as for fp16x2."""
    return 2.0 ** -10 + 2.0 ** -20 + 2.0 * (d + 32) * 2.0 ** -24 + 1e-6


def eps_rel_simt(d):
    """Relative bound on |approx d^2 - exact d^2| / (|x~|^2 + |y~|^2) for the fp32 CUDA-core pass:
    (d + 11) * 2^-24 from a standard rounding analysis (dot product, norms, centring), doubled."""
    return 2.0 * (d + 16) * 2.0 ** -24


def knn_kernel(Xq, ref, qry=None, *, knn, knn_max=None, decay=40, thresh=1e-4, bandwidth=None,
               bandwidth_scale=1.0, S=None, impl=None):
    """Raw (unsymmetrised) kNN / alpha-decay kernel from the rows of ``Xq`` to ``ref`` as DeviceCSR.

    Semantics = reference ``kNNGraph.build_kernel_to_data`` (graphs.py:819-982) with the float64
    path: ``knn`` is the effective neighbour count (callers pass knn+1 for the self-including
    in-sample build), ``knn_max`` likewise.  Returns (csr, info) with info['bandwidth'] (device),
    info['nzero'] (zero-distance candidates per row, for duplicate warnings).
    """
    E.require_cuda()
    if qry is None:
        qry = ref
    nq, nr, d = qry.n, ref.n, ref.d
    if qry.is64 != ref.is64 or qry.metric != ref.metric:
        raise ValueError("query and reference operands must have the same dtype and metric")
    x64 = int(ref.is64) | (METRIC_CODE[ref.metric] << 1)            # x_kind of the refine entry points
    l1 = ref.metric == "cityblock"
    binary = decay is None
    knn = int(min(knn, nr))
    kmax = 0 if knn_max is None else int(knn_max)
    if kmax >= nr:
        kmax = 0
    if l1:
        impl = "simt"
    if impl is None:
        # GTB_SEARCH_IMPL is a preference: a flavour that cannot take this shape (feature count beyond the
        # resident query tile, knn beyond the 2 x 32 candidate lists) hands over to the next one
        want = default_impl()
        order = {"auto": (AUTO_TC,) + tuple(x for x in ("tch", "tc16", "tc") if x != AUTO_TC) + ("simt",),
                 "tc": ("tc", "tc16", "simt"), "tc16": ("tc16", "tc", "simt"), "tch": ("tch", "tc16", "tc", "simt"),
                 "tch1": ("tch1", "tch", "tc16", "tc", "simt"), "simt": ("simt",)}.get(want)
        if order is None:
            raise ValueError("GTB_SEARCH_IMPL must be auto, tc, tc16, tch, tch1 or simt (got %r)" % (want,))
        impl = "simt"
        for cand_impl in order:
            # auto: the one-product flavour pays off only with seeded thresholds, i.e. when the reference set is
            # large enough for a strided sample (cold, its list of 64 costs more than the second product saves:
            # C3 at 50k x 50 6.7 ms against 2.5 ms) -- unless the longer list is what the request needs
            if (cand_impl == "tch1" and want == "auto" and not seedable(ref) and knn + 8 <= 32):
                continue
            # neighbours asked for must fit the candidate lists: 2 x 32 entries, one list of 64 for tch1
            if cand_impl == "simt" or (knn + 8 <= (64 if cand_impl == "tch1" else 32) and S in (None, 32, 64)
                                       and ref.tc_ok(TC_DTYPE[cand_impl])):
                impl = cand_impl
                break
    dev = _dev()
    ntau = 1
    tcd = TC_DTYPE.get(impl, 0)
    tc_scale = 1.0
    if impl in TC_DTYPE:
        if not ref.tc_ok(tcd):
            raise ValueError("tensor-core search: d = {} does not fit the resident query tile".format(d))
        import os
        # two candidate lists per row (one per epilogue group) of `ls` entries each.  Short lists (16) halve the
        # selection work of the sweep; they are used when the neighbourhood asked for is small (knn + 8 <= 16) --
        # rows whose kernel support is wider fail certification and are finished by the radius pass.
        want = max(knn, kmax)                                   # knn_max neighbours must fit the lists as well
        # (fp16x2 runs two query tiles per CTA and keeps ONE list of 32 per row, which covers want + 8 <= 32)
        two_tiles = tcd == 2 and int(os.environ.get("GTB_TC_QTILES", "2")) == 2 and ref.kp(2) <= 128
        short_ok = want + 8 <= (32 if two_tiles else 16)
        ls = int(os.environ.get("GTB_TC_LIST", "16" if short_ok else "32"))
        qtiles = 2 if (two_tiles and ls == 16) else 1
        if tcd == 3:
            ls, qtiles = 32, 2                                  # one product: one list of 64 per row, always two tiles
        if ls not in (16, 32) or knn > ls * qtiles:
            raise ValueError("GTB_TC_LIST must be 16 or 32 and >= knn")
        S, stride, ntau = 2 * ls, 2 * ls, 2
        cluster = min(tc_cluster(), 2) if tcd else tc_cluster()
        if cluster not in (1, 2, 4):
            raise ValueError("GTB_TC_CLUSTER must be 1, 2 or 4")
        # grid-wide pacing of the TMA producers: one caller-owned device word per launch (no library state)
        pace = _empty((1,), torch.int32) if int(os.environ.get("GTB_TC_PACING", "1")) else None
        if knn > (64 if tcd == 3 else 32):
            raise NotImplementedError("knn={} exceeds the tensor-core candidate lists (2 x 32)".format(knn))
        eps_rel = (eps_rel_tc, eps_rel_tc16, eps_rel_tch, eps_rel_tch1)[tcd](d)
        if tcd >= 2:
            tc_scale = fp16_scale(max(qry.norm_max(), ref.norm_max()))
        q_hi, q_lo, q_n2 = qry.tc(0, tcd, tc_scale)
        r_hi, r_lo, _ = ref.tc(1, tcd, tc_scale)
        Kp = ref.kp(tcd)
        cand = _empty((nq, stride), torch.int32)
        tau = _empty((nq, ntau), torch.float32)
        scratch = _empty((E.lib().gtb_tc_scratch_bytes(qry.n_pad),), torch.uint8)
        s2 = tc_scale * tc_scale
        # qtiles == 2: two query tiles per CTA share every reference stage (half the L2 traffic per MMA)
        qn2_s = q_n2 * s2 if tcd >= 2 else q_n2
        seed = None
        stride_s = seed_stride()
        if tcd == 3 and seedable(ref):
            # threshold seeds: the 8th smallest tile minimum over every stride-th reference tile (6 % of the tensor work
            # of a sweep, no selection work); it sits near the (8 stride)-th smallest distance overall, so the full
            # sweep starts with a threshold that admits ~8 stride points per row instead of climbing down from +inf
            # (64 ln(N / 64) threshold updates)
            seed = _empty((nq, ntau), torch.float32)
            E.call("gtb_knn_seed_tc", q_hi, qn2_s, nq, qry.n_pad, r_hi, nr, ref.n_pad, Kp, cluster, stride_s, seed, pace)
        if tcd == 3:
            E.call("gtb_knn_topk_tc_seeded", q_hi, q_lo, qn2_s, nq, qry.n_pad, r_hi, r_lo, nr, ref.n_pad, Kp, tcd, ls,
                   cluster, qtiles, seed, 1, cand, scratch, tau, pace)
        else:
            E.call("gtb_knn_topk_tc", q_hi, q_lo, qn2_s, nq, qry.n_pad, r_hi, r_lo, nr, ref.n_pad, Kp, tcd, ls,
                   cluster, qtiles, cand, scratch, tau, pace)
        del seed
        if tcd >= 2:
            tau = tau / s2                                   # scaled squared distances -> data units (inf stays inf)
        del scratch
    elif impl == "simt":
        if S is None:
            S = choose_S(knn, binary)
        stride = S
        tau = _empty((nq,), torch.float32)
        cand = _empty((nq, S), torch.int32)
        if l1:
            eps_rel = eps_rel_l1(d)
            q_n2 = qry.l1_norms()[0] if ref.is64 else None       # float32 inputs: the search copy is exact
            E.call("gtb_knn_topk_simt_l1", qry.XT, nq, qry.n_pad, ref.XT, nr, ref.n_pad, ref.d_pad, S, cand, tau)
        else:
            eps_rel = eps_rel_simt(d)
            q_n2 = qry.n2
            E.call("gtb_knn_topk_simt", qry.XT, qry.n2, nq, qry.n_pad, ref.XT, ref.n2, nr, ref.n_pad, ref.d_pad, S,
                   cand, tau)
    else:
        raise ValueError("unknown impl %r" % (impl,))

    # bandwidth argument
    bw_mode, bw_fixed = 0, None
    if not binary and bandwidth is not None:
        if np.ndim(bandwidth) == 0:
            bw_mode, bw_fixed = 1, torch.tensor([float(bandwidth)], dtype=torch.float64, device=dev)
        else:
            bw_arr = np.asarray(bandwidth, dtype=np.float64).reshape(-1)
            if bw_arr.shape[0] != nq:
                raise ValueError("bandwidth must be a scalar or have one entry per row ({}), got {}".format(
                    nq, bw_arr.shape[0]))
            bw_mode, bw_fixed = 2, torch.from_numpy(bw_arr).to(dev)
    decay_f = -1.0 if binary else float(decay)
    thresh_f = 1.0 if binary else float(thresh)

    st_idx = _empty((nq, S), torch.int32)
    st_val = _empty((nq, S), torch.float64)
    n_keep = _empty((nq,), torch.int32)
    bw_out = _empty((nq,), torch.float64)
    lim2 = _empty((nq,), torch.float32)
    status = _empty((nq,), torch.int32)
    nzero = _empty((nq,), torch.int32)
    maxnorm = (ref.l1_norms()[1] if ref.is64 else 0.0) if l1 else ref.maxnorm
    E.call("gtb_refine_topk", qry.X, nq, ref.X, d, x64, cand, S, stride, tau, ntau, q_n2, maxnorm, eps_rel, knn, kmax, decay_f,
           thresh_f, bw_fixed, bw_mode, float(bandwidth_scale), st_idx, st_val, n_keep, bw_out, lim2, status, nzero)

    todo_rows = _empty((nq,), torch.int32)
    count = _empty((1,), torch.int32)
    E.call("gtb_compact_todo", status, nq, todo_rows, count)
    nt = int(count.item())                       # host sync #1
    _STATS.update(rows=nq, radius_rows=nt, S=S, radius_pairs=0, impl=impl)

    seg_ptr = seg_idx = seg_val = n_keep_t = None
    if nt > 0:
        todo_rows = todo_rows[:nt].contiguous()
        nt_pad = (nt + 127) // 128 * 128
        lim_t = torch.zeros((nt_pad,), dtype=torch.float32, device=dev)
        lim_t[:nt] = lim2[todo_rows.long()]
        if impl in TC_DTYPE:
            # the radius rows form their own (small) query operand; build it from the gathered rows
            sub = SearchOperand(qry.X[todo_rows.long()].contiguous(), mean=ref.mean, metric=ref.metric)
            s_hi, s_lo, s_n2 = sub.tc(0, tcd, tc_scale)
            r_hi, r_lo, _ = ref.tc(1, tcd, tc_scale)
            if tcd >= 2:
                s_n2 = s_n2 * (tc_scale * tc_scale)
                lim_t = lim_t * (tc_scale * tc_scale)
        else:
            QT = _empty((ref.d_pad, nt_pad), torch.float32)
            qn2 = _empty((nt_pad,), torch.float32)
            E.call("gtb_gather_operand", qry.XT, qry.n_pad, qry.n2, todo_rows, nt, QT, nt_pad, ref.d_pad, qn2)
        capacity = max(1 << 20, nt * 4 * S)
        while True:
            pairs = _empty((capacity, 2), torch.int32)
            counter = _zeros((1,), torch.int64)
            rowcnt = _zeros((nt_pad,), torch.int32)
            if impl in TC_DTYPE:
                # (the one-product flavour hands its uncertified rows to the two-product sweep on the same operands)
                E.call("gtb_knn_radius_tc", s_hi, s_lo, s_n2, lim_t, nt, nt_pad, r_hi, r_lo, nr, ref.n_pad, Kp,
                       min(tcd, 2), cluster, pairs, capacity, counter, rowcnt, pace)
            elif l1:
                E.call("gtb_knn_radius_simt_l1", QT, lim_t, nt, nt_pad, ref.XT, nr, ref.n_pad, ref.d_pad, pairs,
                       capacity, counter, rowcnt)
            else:
                E.call("gtb_knn_radius_simt", QT, qn2, lim_t, nt, nt_pad, ref.XT, ref.n2, nr, ref.n_pad, ref.d_pad,
                       pairs, capacity, counter, rowcnt)
            npairs = int(counter.item())
            if npairs <= capacity:
                break
            capacity = npairs
        _STATS.update(radius_pairs=npairs)
        seg_ptr = exclusive_scan(rowcnt[:nt].contiguous())
        seg_idx = _empty((max(npairs, 1),), torch.int32)
        seg_val = _empty((max(npairs, 1),), torch.float64)
        cursor = _empty((nt,), torch.int32)
        E.call("gtb_scatter_pairs", pairs, npairs, seg_ptr, cursor, nt, seg_idx)
        n_keep_t = _empty((nt,), torch.int32)
        overflow = _empty((1,), torch.int32)
        longest = int(rowcnt.max().item())
        cap = 64
        while cap < longest:
            cap *= 2
        if cap > BALL_CAP:
            raise RowTooLong(
                "a row has {} neighbours inside the kernel radius; the sparse kNN pipeline handles rows of up to {} "
                "(a kernel this dense calls for graphtype='exact', which evaluates every pair)".format(
                    longest, BALL_CAP))
        E.call("gtb_refine_ball", qry.X, todo_rows, status, nt, ref.X, d, x64, seg_ptr, seg_idx, seg_val, knn, kmax,
               decay_f, thresh_f, bw_fixed, bw_mode, float(bandwidth_scale), n_keep_t, n_keep, bw_out, nzero,
               overflow, cap)

    indptr = exclusive_scan(n_keep)
    nnz = int(indptr[-1].item())                 # host sync #2
    out_idx = _empty((nnz,), torch.int32)
    out_val = _empty((nnz,), torch.float64)
    E.call("gtb_csr_gather", st_idx, st_val, n_keep, status, indptr, nq, S, todo_rows if nt else None, nt,
           seg_ptr, seg_idx, seg_val, n_keep_t, out_idx, out_val)
    _STATS.update(nnz_raw=nnz)
    csr = DeviceCSR(indptr, out_idx, out_val, (nq, nr))
    return csr, {"bandwidth": bw_out, "nzero": nzero}


def sort_rows(ptr, idx, val, n):
    """Order every row of a CSR by column, in place (csrc/symm.cu: tiered segmented sort)."""
    if n == 0 or idx.shape[0] == 0:
        return
    E.call("gtb_csr_sort_rows", ptr, idx, val, n, _empty((1,), torch.int32))


def cursor32(ptr):
    """32-bit copy of the first n entries of a row-pointer array: the per-row scatter cursors."""
    n = ptr.shape[0] - 1
    cur = _empty((max(n, 1),), torch.int32)
    if n:
        E.call("gtb_cast_indptr", ptr, n, cur)
    return cur


def transpose_records(R, sort_min_total=None, pa=None):
    """Rows of R^T as 16-byte edge records {int32 i, int32 j, float64 w} (csrc/symm.cu): histogram of the columns ->
    scan -> scatter (one atomic + one 16-byte store per edge).  Rows come out in arrival order; rows whose length plus
    the matching row of ``pa`` exceeds ``sort_min_total`` are then ordered by column (None: no sort, 0: every row).
    Returns (ptr_t int64 [n_cols + 1], records [nnz, 2] int64)."""
    n_rows, n_cols = R.shape
    cnt = _empty((n_cols,), torch.int32)
    E.call("gtb_transpose_count", R.indices, R.nnz, 0, cnt, n_cols)
    ptr_t = exclusive_scan(cnt)
    rec = _empty((R.nnz, 2), torch.int64)
    if R.nnz:
        E.call("gtb_transpose_scatter", R.indptr, R.indices, R.data, n_rows, 0, 0, cursor32(ptr_t), rec)
        if sort_min_total is not None:
            sort_records(ptr_t, rec, n_cols, pa, sort_min_total)
    return ptr_t, rec


def sort_records(ptr_t, rec, n, pa=None, min_total=0):
    if n and rec.shape[0]:
        E.call("gtb_rec_sort_rows", ptr_t, rec, n, pa, int(min_total), _empty((1,), torch.int32))


def transpose_csr(R):
    """R^T as a DeviceCSR with column-sorted rows."""
    ptr_t, rec = transpose_records(R, sort_min_total=0)
    idx = rec.view(torch.int32).view(-1, 4)[:, 0].contiguous()
    val = rec.view(torch.float64).view(-1, 2)[:, 1].contiguous()
    return DeviceCSR(ptr_t, idx, val, (R.shape[1], R.shape[0]))


def merge_with_transpose(A_ptr, A_idx, A_val, T_ptr, T_rec, n_rows, row0, mode, theta, want_p=True, flags=None):
    """sym(A, T) row by row (csrc/symm.cu sym_merge): returns (outptr, k_idx, k_val, p_val | None, degree, newlen).
    ``T_rec``: record rows of the transposed matrix in any order (long rows are sorted in place by the count call)."""
    newlen = _empty((n_rows,), torch.int32)
    worklist = _empty((n_rows + 1,), torch.int32)        # rows too long for the register path, sorted inside the call
    E.call("gtb_sym_merge_count", A_ptr, A_idx, A_val, T_ptr, T_rec, n_rows, mode, theta, newlen, worklist)
    outptr = exclusive_scan(newlen)
    nnz = int(outptr[-1].item())
    k_idx = _empty((nnz,), torch.int32)
    k_val = _empty((nnz,), torch.float64)
    p_val = _empty((nnz,), torch.float64) if want_p else None
    degree = _empty((n_rows,), torch.float64)
    E.call("gtb_sym_merge_fill", A_ptr, A_idx, A_val, T_ptr, T_rec, n_rows, row0, mode, theta, outptr,
           k_idx, k_val, p_val, degree, flags)
    return outptr, k_idx, k_val, p_val, degree, newlen


def symmetrize_normalize(R, kernel_symm="+", theta=None, anisotropy=0.0, want_p=True):
    """``BaseGraph._build_kernel`` post-processing (base.py:534-592) + ``P`` (base.py:645).

    Returns (K DeviceCSR, P values tensor | None, degree tensor, flags int) where flags bit 0 =
    asymmetric beyond 1e-5 (only evaluated for kernel_symm=None), bit 1 = a row has no diagonal.
    """
    n = R.shape[0]
    square = R.shape[0] == R.shape[1]
    mode = SYM_MODES[kernel_symm]
    flags = _zeros((1,), torch.int32)
    if mode != 3 and not square:
        raise ValueError("symmetrisation needs a square kernel")
    if mode == 3:
        if square:
            E.call("gtb_asym_check", R.indptr, R.indices, R.data, n, flags)
        K = R
        P = _empty((R.nnz,), torch.float64) if want_p else None
        degree = _empty((n,), torch.float64)
        E.call("gtb_row_finalize", K.indptr, K.indices, K.data, n, P, degree, flags, int(square))
    else:
        th = 0.0 if theta is None else float(theta)
        ptr_t, rec = transpose_records(R)                # arrival order: the merge sorts the (few) long rows itself
        outptr, k_idx, k_val, P, degree, _ = merge_with_transpose(
            R.indptr, R.indices, R.data, ptr_t, rec, n, 0, mode, th,
            want_p=want_p and anisotropy == 0, flags=flags)
        K = DeviceCSR(outptr, k_idx, k_val, R.shape)
    if anisotropy != 0:
        E.call("gtb_anisotropy", K.indptr, K.indices, K.data, degree, float(anisotropy), n)
        P = _empty((K.nnz,), torch.float64) if want_p else None
        E.call("gtb_row_finalize", K.indptr, K.indices, K.data, n, P, degree, flags, 0)
    _STATS.update(nnz_sym=K.nnz)
    return K, P, degree, int(flags.item())


def row_normalize(K):
    """sklearn ``normalize(K, 'l1', axis=1)`` on a DeviceCSR -> values tensor."""
    P = _empty((K.nnz,), torch.float64)
    flags = _zeros((1,), torch.int32)
    if K.shape[0]:
        E.call("gtb_row_finalize", K.indptr, K.indices, K.data, K.shape[0], P, None, flags, 0)
    return P


def row_sums(K):
    """Row L1 sums of a DeviceCSR (``np.sum(K, 1)`` for the non-negative kernels of this package)."""
    deg = _empty((K.shape[0],), torch.float64)
    flags = _zeros((1,), torch.int32)
    E.call("gtb_row_finalize", K.indptr, K.indices, K.data, K.shape[0], None, deg, flags, 0)
    return deg


def csr_from_scipy(M):
    """scipy sparse -> DeviceCSR (canonical CSR: sorted columns, duplicates summed)."""
    from scipy import sparse
    M = sparse.csr_matrix(M, dtype=np.float64)
    if not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    dev = _dev()
    return DeviceCSR(torch.from_numpy(M.indptr.astype(np.int64)).to(dev),
                     torch.from_numpy(M.indices.astype(np.int32)).to(dev),
                     torch.from_numpy(np.ascontiguousarray(M.data)).to(dev), M.shape)


def spmm(A, B, vals=None):
    """``A.dot(B)`` on the device (csrc/spmm.cu): A = DeviceCSR (``vals`` substitutes its values, e.g. the
    diffusion operator sharing K's structure), B = float64 CUDA tensor [n_cols, f] (row-major, any row
    stride).  Accumulation order and rounding are scipy's, so the result is bit-identical to the host product
    the reference computes (base.py:1229)."""
    if B.dim() != 2 or B.shape[0] != A.shape[1]:
        raise ValueError("shapes {} and {} not aligned".format(A.shape, tuple(B.shape)))
    if B.dtype != torch.float64 or B.stride(1) != 1:
        B = B.to(torch.float64).contiguous()
    f = B.shape[1]
    out = _empty((A.shape[0], f), torch.float64)
    if A.shape[0] == 0 or f == 0:
        return out
    E.call("gtb_spmm_csr", A.indptr, A.indices, A.data if vals is None else vals, A.shape[0], B, B.stride(0), f,
           out, f)
    return out
