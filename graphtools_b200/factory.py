"""``graphtools_b200.Graph`` -- the drop-in for ``graphtools.Graph`` (reference graphtools/api.py:14-295).

Same signature, same automatic graph-type selection, same error messages; the returned object is an
instance of ``{kNN,MNN,Traditional}[Landmark]Graph`` from this package, whose heavy lifting runs on
the B200 engine.
"""
import warnings

import numpy as np

from . import graphs
from .logging_util import logger as _logger


def Graph(data, n_pca=None, rank_threshold=None, knn=5, decay=40, bandwidth=None, bandwidth_scale=1.0,
          knn_max=None, anisotropy=0, distance="euclidean", thresh=1e-4, kernel_symm="+", theta=None,
          precomputed=None, beta=1, sample_idx=None, adaptive_k=None, n_landmark=None, n_svd=100,
          random_landmarking=False, n_jobs=-1, verbose=False, random_state=None, graphtype="auto",
          use_pygsp=False, initialize=True, **kwargs):
    """Create a graph built on data; see the reference docstring (api.py:43-184) for parameters."""
    _logger.set_level(verbose)
    if sample_idx is not None and len(np.unique(sample_idx)) == 1:
        warnings.warn("Only one unique sample. Not using MNNGraph")
        sample_idx = None
        if graphtype == "mnn":
            graphtype = "auto"
    if graphtype == "auto":
        if sample_idx is not None:
            graphtype = "mnn"            # only mnn does batch correction
        elif precomputed is not None:
            graphtype = "exact"          # precomputed requires the exact graph
        elif decay is None:
            graphtype = "knn"
        elif (thresh == 0 and knn_max is None) or callable(bandwidth):
            graphtype = "exact"          # full distance matrix needed
        else:
            graphtype = "knn"            # decay kernel with a threshold: sparse kNN search

    if graphtype == "knn":
        base = graphs.kNNGraph
        if precomputed is not None:
            raise ValueError("kNNGraph does not support precomputed values. Use `graphtype='exact'` or "
                             "`precomputed=None`")
        if sample_idx is not None:
            raise ValueError("kNNGraph does not support batch correction. Use `graphtype='mnn'` or "
                             "`sample_idx=None`")
    elif graphtype == "mnn":
        base = graphs.MNNGraph
        if precomputed is not None:
            raise ValueError("MNNGraph does not support precomputed values. Use `graphtype='exact'` and "
                             "`sample_idx=None` or `precomputed=None`")
    elif graphtype == "exact":
        base = graphs.TraditionalGraph
        if sample_idx is not None:
            raise ValueError("TraditionalGraph does not support batch correction. Use `graphtype='mnn'` or "
                             "`sample_idx=None`")
    else:
        raise ValueError("graphtype '{}' not recognized. Choose from ['knn', 'mnn', 'exact', 'auto']"
                         .format(graphtype))

    parents = [base]
    msg = "Building {} graph".format(graphtype)
    if n_landmark is not None:
        parents.append(graphs.LandmarkGraph)
        msg += " with landmarks"
    if use_pygsp:
        parents.append(graphs.PyGSPGraph)
        msg += " with PyGSP inheritance" if len(parents) > 2 else " and PyGSP inheritance"
    _logger.log_debug(msg)

    cls_name = "".join(p.__name__.replace("Graph", "") for p in parents) + "Graph"
    try:
        cls = getattr(graphs, cls_name)
    except AttributeError:
        raise RuntimeError("unknown graph classes {}".format(parents))

    available = dict(locals())
    params = kwargs
    for parent in parents:
        for name in parent._get_param_names():
            if name in available and name not in ("kwargs", "params", "available"):
                params[name] = available[name]
    _logger.log_debug("Initializing {} with arguments {}".format(
        parents, ", ".join("{}='{}'".format(k, v) for k, v in params.items() if k != "data")))
    return cls(**params)


def from_igraph(G, attribute="weight", **kwargs):
    """igraph.Graph -> TraditionalGraph on its (weighted) adjacency matrix (reference api.py:298-336)."""
    from scipy import sparse
    if "precomputed" in kwargs:
        if kwargs["precomputed"] != "adjacency":
            warnings.warn("Cannot build graph from igraph with precomputed={}. Use 'adjacency' instead.".format(
                kwargs["precomputed"]), UserWarning)
        del kwargs["precomputed"]
    try:
        K = G.get_adjacency(attribute=attribute).data
    except ValueError as e:
        if str(e) == "Attribute does not exist":
            warnings.warn("Edge attribute {} not found. Returning unweighted graph".format(attribute), UserWarning)
        K = G.get_adjacency(attribute=None).data
    return Graph(sparse.coo_matrix(K), precomputed="adjacency", **kwargs)


def read_pickle(path):
    """Load a pickled graph (or any object) from ``path`` (reference api.py:339-354).  Graphs are pickled with their
    host-side results (``BaseGraph.__getstate__`` materialises K / P and drops the device handles)."""
    import pickle
    from .core import BaseGraph
    with open(path, "rb") as f:
        G = pickle.load(f)
    if not isinstance(G, BaseGraph):
        warnings.warn("Returning object that is not a graphtools.base.BaseGraph")
    return G
