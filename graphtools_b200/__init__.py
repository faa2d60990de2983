"""graphtools_b200 -- B200-native graph construction behind the graphtools API.

    import graphtools_b200 as graphtools
    G = graphtools.Graph(X, knn=5, decay=40)
    G.kernel, G.diff_op
"""
from .factory import Graph, from_igraph, read_pickle
from . import graphs
from .graphs import (kNNGraph, TraditionalGraph, MNNGraph, LandmarkGraph, kNNLandmarkGraph, MNNLandmarkGraph,
                     TraditionalLandmarkGraph)

__version__ = "0.1.0"
__all__ = ["Graph", "from_igraph", "read_pickle", "graphs", "kNNGraph", "TraditionalGraph", "MNNGraph", "LandmarkGraph", "kNNLandmarkGraph",
           "MNNLandmarkGraph", "TraditionalLandmarkGraph"]
