import numpy as np
from scipy import sparse


class Graph:
    """Holds W/N/dw/d the way pygsp.graphs.Graph exposes them (enough for the
    reference's own tests to compare graphs)."""

    def __init__(self, W, lap_type="combinatorial", coords=None, plotting=None, **kwargs):
        self.W = sparse.lil_matrix(W) if not sparse.issparse(W) else W
        self.N = W.shape[0]
        self.lap_type = lap_type
        self.coords = coords
        self.plotting = plotting or {}
        Wc = sparse.csr_matrix(W)
        self.dw = np.asarray(Wc.sum(axis=1)).ravel()
        self.d = np.asarray((Wc != 0).sum(axis=1)).ravel()
