"""Concrete graph classes (names as in reference graphtools/graphs.py:562-2002)."""
from .core import Base
from .knn import kNNGraph
from .landmark import LandmarkGraph
from .exact import TraditionalGraph
from .mnn import MNNGraph


class PyGSPGraph(Base):
    """Placeholder for the reference's optional PyGSP bridge (base.py:991-1043): out of scope
    (third-party toolbox, not on the accelerated path) -- selecting it raises."""

    def __init__(self, **kwargs):
        raise NotImplementedError("use_pygsp=True is not supported by graphtools_b200 (pygsp bridge is out of scope)")


class kNNLandmarkGraph(kNNGraph, LandmarkGraph):
    pass


class MNNLandmarkGraph(MNNGraph, LandmarkGraph):
    pass


class TraditionalLandmarkGraph(TraditionalGraph, LandmarkGraph):
    pass


class kNNPyGSPGraph(kNNGraph, PyGSPGraph):
    pass


class MNNPyGSPGraph(MNNGraph, PyGSPGraph):
    pass


class TraditionalPyGSPGraph(TraditionalGraph, PyGSPGraph):
    pass


class kNNLandmarkPyGSPGraph(kNNGraph, LandmarkGraph, PyGSPGraph):
    pass


class MNNLandmarkPyGSPGraph(MNNGraph, LandmarkGraph, PyGSPGraph):
    pass


class TraditionalLandmarkPyGSPGraph(TraditionalGraph, LandmarkGraph, PyGSPGraph):
    pass
