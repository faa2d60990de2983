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
        if distance != "euclidean":
            raise NotImplementedError(
                "graphtools_b200 accelerates the Euclidean metric only (got distance={!r})".format(distance))
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
        try:
            return self._ref_operand
        except AttributeError:
            self._ref_operand = pipeline.SearchOperand(self._dense_f32(self.data_nu))
            return self._ref_operand

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

    def _build_kernel_sharded(self, ref, knn_max):
        """One process per GPU: this rank builds the kernel rows of its contiguous query shard against the
        replicated reference set, then the raw CSR shards are all-gathered (NCCL) so that symmetrisation
        and normalisation are local.  Every rank ends up with the complete, identical kernel."""
        import torch
        import torch.distributed as dist
        from . import distributed as gd
        world, rank = dist.get_world_size(), dist.get_rank()
        n = ref.n
        bounds = [gd.shard_bounds(n, world, r) for r in range(world)]
        lo, hi = bounds[rank]
        if hi > lo:
            qry = pipeline.SearchOperand(ref.X[lo:hi], mean=ref.mean)
            Rl, _ = self._kernel_device(qry, ref, knn=self.knn + 1, knn_max=knn_max, bandwidth=self.bandwidth,
                                        bandwidth_scale=self.bandwidth_scale)
            row_len = (Rl.indptr[1:] - Rl.indptr[:-1]).to(torch.int32)
            idx, val = Rl.indices, Rl.data
        else:
            row_len = torch.zeros((0,), dtype=torch.int32, device=ref.X.device)
            idx = torch.zeros((0,), dtype=torch.int32, device=ref.X.device)
            val = torch.zeros((0,), dtype=torch.float64, device=ref.X.device)
        indptr, idx, val = gd.allgather_csr_rows(row_len, idx, val, [b[1] - b[0] for b in bounds],
                                                 pipeline.exclusive_scan)
        return pipeline.DeviceCSR(indptr, idx, val, (n, n))

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
        with _logger.log_task("KNN search"):
            qry = pipeline.SearchOperand(self._dense_f32(Y), mean=ref.mean)
            R, info = self._kernel_device(qry, ref, knn=knn, knn_max=knn_max, bandwidth=bandwidth,
                                          bandwidth_scale=bandwidth_scale)
        self._check_duplicates(info, qry, ref)
        return R

    def build_kernel_to_data(self, Y, knn=None, knn_max=None, bandwidth=None, bandwidth_scale=None):
        """Kernel from new points ``Y`` to ``self.data`` as scipy CSR [n_y, n] (graphs.py:819-982)."""
        return self._kernel_to_data_device(Y, knn, knn_max, bandwidth, bandwidth_scale).to_scipy()
