"""MNNGraph stub (filled in next)."""
from .core import DataGraph


class MNNGraph(DataGraph):
    def __init__(self, data, sample_idx, knn=5, beta=1, n_pca=None, decay=None, adaptive_k=None, bandwidth=None,
                 distance="euclidean", thresh=1e-4, n_jobs=1, **kwargs):
        raise NotImplementedError("MNNGraph: device path under construction")
