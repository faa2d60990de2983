"""kNNGraph on the CUDA engine (reference graphtools/graphs.py:562-982).

Constructor arguments, validation, warnings and ``get_params``/``set_params`` follow the reference;
``build_kernel`` / ``build_kernel_to_data`` run the fused distance/top-k + float64 refine pipeline
(pipeline.knn_kernel) instead of sklearn ``NearestNeighbors`` + the Python CSR loop.
"""
import warnings

import numpy as np
from scipy import sparse

from . import pipeline
from .core import DataGraph
from .logging_util import logger as _logger


class kNNGraph(DataGraph):
    """K nearest neighbours graph, optionally with alpha-decay affinities."""

    def __init__(self, data, knn=5, decay=None, knn_max=None, search_multiplier=6, bandwidth=None,
                 bandwidth_scale=1.0, distance="euclidean", thresh=1e-4, n_pca=None, **kwargs):
        if decay is not None:
            if thresh <= 0 and knn_max is None:
                raise ValueError("Cannot instantiate a kNNGraph with `decay=None`, "
                                 "`thresh=0` and `knn_max=None`. Use a TraditionalGraph instead.")
            elif thresh < np.finfo(float).eps:
                thresh = np.finfo(float).eps
        if callable(bandwidth):
            raise NotImplementedError("Callable bandwidth is only supported by"
                                      " graphtools.graphs.TraditionalGraph.")
        if knn is None and bandwidth is None:
            raise ValueError("Either `knn` or `bandwidth` must be provided.")
        elif knn is None and bandwidth is not None:
            knn = 5  # the implementation needs some knn value
        if decay is None and bandwidth is not None:
            warnings.warn("`bandwidth` is not used when `decay=None`.", UserWarning)
        if knn > data.shape[0] - 2:
            warnings.warn("Cannot set knn ({k}) to be greater than "
                          "n_samples - 2 ({n}). Setting knn={n}".format(k=knn, n=data.shape[0] - 2))
            knn = data.shape[0] - 2
        if knn_max is not None and knn_max < knn:
            warnings.warn("Cannot set knn_max ({knn_max}) to be less than "
                          "knn ({knn}). Setting knn_max={knn}".format(knn=knn, knn_max=knn_max))
            knn_max = knn
        if n_pca in [None, 0, False] and data.shape[1] > 500:
            warnings.warn("Building a kNNGraph on data of shape {} is "
                          "expensive. Consider setting n_pca.".format(data.shape), UserWarning)
        if distance not in ("euclidean", "cosine", "cityblock", "manhattan", "l1"):
            raise NotImplementedError(
                "graphtools_b200 accelerates the euclidean, cosine and cityblock metrics (got distance={!r})".format(
                    distance))
        self.knn = knn
        self.knn_max = knn_max
        self.search_multiplier = search_multiplier
        self.decay = decay
        self.bandwidth = bandwidth
        self.bandwidth_scale = bandwidth_scale
        self.distance = distance
        self.thresh = thresh
        super().__init__(data, n_pca=n_pca, **kwargs)

    def get_params(self):
        params = super().get_params()
        params.update({"knn": self.knn, "decay": self.decay, "bandwidth": self.bandwidth,
                       "bandwidth_scale": self.bandwidth_scale, "knn_max": self.knn_max,
                       "distance": self.distance, "thresh": self.thresh, "n_jobs": self.n_jobs,
                       "random_state": self.random_state, "verbose": self.verbose})
        return params

    def set_params(self, **params):
        if "knn" in params and params["knn"] != self.knn:
            raise ValueError("Cannot update knn. Please create a new graph")
        # the reference compares knn_max with self.knn (graphs.py:721); kept as is
        if "knn_max" in params and params["knn_max"] != self.knn:
            raise ValueError("Cannot update knn_max. Please create a new graph")
        for name in ("decay", "bandwidth", "bandwidth_scale", "distance"):
            if name in params and params[name] != getattr(self, name):
                raise ValueError("Cannot update {}. Please create a new graph".format(name))
        if "thresh" in params and params["thresh"] != self.thresh and self.decay != 0:
            raise ValueError("Cannot update thresh. Please create a new graph")
        if "n_jobs" in params:
            self.n_jobs = params["n_jobs"]
        if "random_state" in params:
            self.random_state = params["random_state"]
        if "verbose" in params:
            self.verbose = params["verbose"]
        super().set_params(**params)
        return self

    # ------------------------------------------------------------------ device side
    @property
    def knn_tree(self):
        """The fitted search structure: the device-resident search operand of ``data_nu``
        (centred k-major copy + norms).  Stands in for the reference's sklearn NearestNeighbors
        (graphs.py:748-769)."""
        if "_ref_operand" not in self.__dict__:
            self._ref_operand = pipeline.SearchOperand(self._reference_rows(), metric=self.distance)
        return self._ref_operand

    def _reference_rows(self):
        """``data_nu`` in HBM.  One process per GPU: every rank uploads its own row block of the host array and the
        blocks are assembled by an NCCL all-gather (collective 1) instead of every rank copying all
        of X across its PCIe link."""
        import torch
        from . import distributed as gd
        A = self.data_nu
        host_dense = isinstance(A, np.ndarray) or (isinstance(A, torch.Tensor) and not A.is_cuda)
        if (gd.active() and host_dense and getattr(self, "_dev_data_nu", None) is None
                and A.shape[0] >= gd.MIN_ROWS_PER_RANK * gd.world_size()):
            is64 = (A.dtype == np.float64) if isinstance(A, np.ndarray) else (A.dtype == torch.float64)
            return gd.upload_sharded(A, torch.float64 if is64 else torch.float32, pipeline._dev())
        return self._dense_f32(A)

    def build_kernel(self):
        """Raw in-sample kernel: every sample queried against all samples with knn+1 neighbours
        (self included), graphs.py:771-785."""
        knn_max = self.knn_max + 1 if self.knn_max else None
        from . import distributed as gd
        with _logger.log_task("KNN search"):
            ref = self.knn_tree
            if gd.active() and np.ndim(self.bandwidth) == 0:
                return self._build_kernel_sharded(ref, knn_max)
            R, info = self._kernel_device(ref, ref, knn=self.knn + 1, knn_max=knn_max,
                                          bandwidth=self.bandwidth, bandwidth_scale=self.bandwidth_scale)
        self._check_duplicates(info, ref, ref)
        self._dev_bandwidth = info["bandwidth"]
        return R

    def _local_raw_rows(self, ref, knn_max, Xq=None, knn=None, bandwidth=None, bandwidth_scale=None):
        """(bounds, indptr, row_len int32 [m], indices, data, info) of this rank's raw kernel rows [lo, hi) of the
        query set ``Xq`` (default: the reference set itself, i.e. the in-sample build)."""
        import torch
        import torch.distributed as dist
        from . import distributed as gd
        world, rank = dist.get_world_size(), dist.get_rank()
        Xq = ref.X if Xq is None else Xq
        knn = self.knn + 1 if knn is None else knn
        bandwidth = self.bandwidth if bandwidth is None else bandwidth
        bandwidth_scale = self.bandwidth_scale if bandwidth_scale is None else bandwidth_scale
        bounds = [gd.shard_bounds(Xq.shape[0], world, r) for r in range(world)]
        lo, hi = bounds[rank]
        dev = ref.X.device
        if hi > lo:
            qry = pipeline.SearchOperand(Xq[lo:hi], mean=ref.mean, metric=ref.metric)
            Rl, info = self._kernel_device(qry, ref, knn=knn, knn_max=knn_max, bandwidth=bandwidth,
                                           bandwidth_scale=bandwidth_scale)
            return (bounds, Rl.indptr, (Rl.indptr[1:] - Rl.indptr[:-1]).to(torch.int32), Rl.indices, Rl.data,
                    info)
        z = torch.zeros((0,), dtype=torch.int32, device=dev)
        return (bounds, torch.zeros((1,), dtype=torch.int64, device=dev), z, z,
                torch.zeros((0,), dtype=torch.float64, device=dev), {"nzero": z, "bandwidth": None})

    def _build_kernel_sharded(self, ref, knn_max):
        """One process per GPU, raw kernel only (used when the symmetrisation cannot be sharded): this rank
        builds the rows of its contiguous query shard against the replicated reference set, the raw CSR
        shards are all-gathered (NCCL) and every rank continues with the complete matrix."""
        from . import distributed as gd
        bounds, _, row_len, idx, val, _ = self._local_raw_rows(ref, knn_max)
        n = ref.n
        indptr, idx, val = gd.allgather_csr_rows(row_len, idx, val, [b[1] - b[0] for b in bounds],
                                                 pipeline.exclusive_scan)
        return pipeline.DeviceCSR(indptr, idx, val, (n, n))

    def _build_kernel(self):
        """Sharded build when torch.distributed is initialised: rows are sharded, every raw edge
        (i, j, w) is routed to the owner of column j with one NCCL all-to-all of packed records (bucketed by kernels,
        csrc/symm.cu route_count / route_fill), each rank turns what it received into the rows of the transposed matrix
        (records_count / scan / records_scatter / csr_sort_rows), merges them with its raw rows and normalises its
        shard with the same kernel as the single-GPU build (sym_merge) -- bit-identical for any rank count.  The
        result stays row-sharded in HBM (``_dev_shard``); the complete K / P are assembled on demand
        (core.BaseGraph._gather_shards on the device, _materialize_shards on the host)."""
        from . import distributed as gd
        sharded = (gd.active() and np.ndim(self.bandwidth) == 0 and self.anisotropy == 0
                   and self.kernel_symm in ("+", "*", "mnn"))
        if not sharded:
            return super()._build_kernel()
        import torch
        import torch.distributed as dist
        from . import _engine as E
        knn_max = self.knn_max + 1 if self.knn_max else None
        ref = self.knn_tree
        n = ref.n
        with _logger.log_task("KNN search"):
            bounds, indptr_a, row_len, idx, val, info = self._local_raw_rows(ref, knn_max)
        rank = dist.get_rank()
        lo, hi = bounds[rank]
        m = hi - lo
        dev = idx.device
        rec = gd.exchange_edges(indptr_a, idx, val, lo, bounds)
        mode = pipeline.SYM_MODES[self.kernel_symm]
        theta = 0.0 if self.theta is None else float(self.theta)
        flags = pipeline._zeros((1,), torch.int32)
        if m > 0:
            k = rec.shape[0]
            cnt = pipeline._empty((m,), torch.int32)
            E.call("gtb_records_count", rec, k, lo, cnt, m)
            ptr_t = pipeline.exclusive_scan(cnt)
            t_rec = pipeline._empty((k, 2), torch.int64)
            E.call("gtb_records_scatter", rec, k, lo, pipeline.cursor32(ptr_t), t_rec)
            outptr, k_idx, k_val, p_val, deg, newlen = pipeline.merge_with_transpose(
                indptr_a, idx, val, ptr_t, t_rec, m, lo, mode, theta, want_p=True, flags=flags)
        else:
            newlen = torch.zeros((0,), dtype=torch.int32, device=dev)
            outptr = torch.zeros((1,), dtype=torch.int64, device=dev)
            k_idx = torch.zeros((0,), dtype=torch.int32, device=dev)
            k_val = p_val = deg = torch.zeros((0,), dtype=torch.float64, device=dev)
        self._dev_shard = {"bounds": bounds, "n": n, "lo": lo, "hi": hi, "indptr": outptr, "row_len": newlen,
                           "indices": k_idx, "data": k_val, "P": p_val, "degree": deg, "flags": flags}
        if info.get("bandwidth") is not None:
            self._dev_bandwidth_shard = info["bandwidth"]
        pipeline._STATS.update(nnz_sym_local=int(k_idx.shape[0]))
        return None

    def _kernel_device(self, qry, ref, knn, knn_max, bandwidth, bandwidth_scale):
        if self.decay is None or self.thresh == 1:
            return pipeline.knn_kernel(None, ref, qry, knn=knn, knn_max=None, decay=None)
        return pipeline.knn_kernel(None, ref, qry, knn=knn, knn_max=knn_max, decay=self.decay,
                                   thresh=self.thresh, bandwidth=bandwidth, bandwidth_scale=bandwidth_scale)

    def _check_duplicates(self, info, qry, ref):
        """Warn about zero distances between distinct samples (graphs.py:787-817)."""
        if self.decay is None or self.thresh == 1:
            return
        nzero = info["nzero"]
        n_dup_rows = int((nzero > 1).sum().item())
        if n_dup_rows == 0:
            return
        total = int((nzero - 1).clamp(min=0).sum().item())
        if total < 20 and qry is ref:
            rows = (nzero > 1).nonzero().flatten().cpu().numpy()
            Xh = ref.X.cpu().numpy()
            pairs = set()
            for i in rows:
                same = np.flatnonzero((Xh == Xh[i]).all(axis=1))
                for j in same:
                    if j < i:
                        pairs.add((int(j), int(i)))
            names = ", ".join("{} and {}".format(a, b) for a, b in sorted(pairs))
            warnings.warn("Detected zero distance between samples {}. Consider removing duplicates to avoid "
                          "errors in downstream processing.".format(names), RuntimeWarning)
        else:
            warnings.warn("Detected zero distance between {} pairs of samples. Consider removing duplicates to "
                          "avoid errors in downstream processing.".format(total // 2), RuntimeWarning)

    def _kernel_to_data_device(self, Y, knn=None, knn_max=None, bandwidth=None, bandwidth_scale=None):
        if knn is None:
            knn = self.knn
        if bandwidth is None:
            bandwidth = self.bandwidth
        if bandwidth_scale is None:
            bandwidth_scale = self.bandwidth_scale
        if knn > self.data.shape[0]:
            warnings.warn("Cannot set knn ({k}) to be greater than "
                          "n_samples ({n}). Setting knn={n}".format(k=knn, n=self.data_nu.shape[0]))
            knn = self.data_nu.shape[0]
        Y = self._check_extension_shape(Y)
        ref = self.knn_tree
        from . import distributed as gd
        with _logger.log_task("KNN search"):
            Yd = self._dense_f32(Y).to(ref.X.dtype)
            if gd.active() and np.ndim(bandwidth) == 0 and Yd.shape[0] >= gd.MIN_ROWS_PER_RANK * gd.world_size():
                return self._kernel_to_data_sharded(Yd, ref, knn, knn_max, bandwidth, bandwidth_scale)
            qry = pipeline.SearchOperand(Yd, mean=ref.mean, metric=ref.metric)
            R, info = self._kernel_device(qry, ref, knn=knn, knn_max=knn_max, bandwidth=bandwidth,
                                          bandwidth_scale=bandwidth_scale)
        self._check_duplicates(info, qry, ref)
        return R

    def _kernel_to_data_sharded(self, Yd, ref, knn, knn_max, bandwidth, bandwidth_scale):
        """Out-of-sample kernel with the query rows sharded over the ranks (MNN cross-batch blocks,
        ``extend_to_data``): each rank searches its contiguous block of ``Y`` against the replicated reference
        set, the CSR row shards are all-gathered (NCCL) and every rank returns the complete [n_y, n] matrix --
        bit-identical to the single-GPU result because every row still sees the whole reference set."""
        from . import distributed as gd
        bounds, _, row_len, idx, val, info = self._local_raw_rows(ref, knn_max, Xq=Yd, knn=knn, bandwidth=bandwidth,
                                                                  bandwidth_scale=bandwidth_scale)
        heights = [b[1] - b[0] for b in bounds]
        indptr, idx, val = gd.allgather_csr_rows(row_len, idx, val, heights, pipeline.exclusive_scan)
        if not (self.decay is None or self.thresh == 1):
            nzero = gd._allgather_padded(info["nzero"], heights)
            self._check_duplicates({"nzero": nzero}, None, ref)
        return pipeline.DeviceCSR(indptr, idx, val, (Yd.shape[0], ref.n))

    def build_kernel_to_data(self, Y, knn=None, knn_max=None, bandwidth=None, bandwidth_scale=None):
        """Kernel from new points ``Y`` to ``self.data`` as scipy CSR [n_y, n] (graphs.py:819-982)."""
        return self._kernel_to_data_device(Y, knn, knn_max, bandwidth, bandwidth_scale).to_scipy()
