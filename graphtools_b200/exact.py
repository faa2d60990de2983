"""TraditionalGraph stub (filled in next)."""
from .core import DataGraph


class TraditionalGraph(DataGraph):
    def __init__(self, data, knn=5, decay=40, bandwidth=None, bandwidth_scale=1.0, distance="euclidean",
                 n_pca=None, thresh=1e-4, precomputed=None, **kwargs):
        raise NotImplementedError("TraditionalGraph: device path under construction")
