"""LandmarkGraph stub (filled in next)."""
from .core import DataGraph


class LandmarkGraph(DataGraph):
    def __init__(self, data, n_landmark=2000, n_svd=100, random_landmarking=False, **kwargs):
        raise NotImplementedError("LandmarkGraph: device path under construction")
