"""Host-side contract of the drop-in API (no GPU needed: graphs are created with initialize=False, so only
argument handling, graph-type selection, error text and warnings run).  The expected strings are the ones the
reference's own tests assert on (reference test/test_knn.py:29-114, test_exact.py:24-123, test_mnn.py:23-326,
test_landmark.py:24-40, test_api.py:150-165)."""
import warnings

import numpy as np
import pytest

import graphtools_b200 as gt

rng = np.random.default_rng(0)
DATA = rng.normal(size=(60, 8)).astype(np.float32)


def build(data=DATA, **kw):
    kw.setdefault("initialize", False)
    kw.setdefault("verbose", 0)
    return gt.Graph(data, **kw)


@pytest.mark.parametrize("kw,cls", [
    (dict(), "kNNGraph"),
    (dict(decay=None), "kNNGraph"),
    (dict(thresh=0), "TraditionalGraph"),
    (dict(thresh=0, knn_max=10), "kNNGraph"),
    (dict(bandwidth=lambda d: 1.0), "TraditionalGraph"),
    (dict(graphtype="exact"), "TraditionalGraph"),
    (dict(sample_idx=np.arange(60) % 3, kernel_symm="mnn", theta=0.5), "MNNGraph"),
    (dict(n_landmark=20), "kNNLandmarkGraph"),
    (dict(n_landmark=20, graphtype="exact"), "TraditionalLandmarkGraph"),
    (dict(n_landmark=20, sample_idx=np.arange(60) % 2), "MNNLandmarkGraph"),
])
def test_graph_type_selection(kw, cls):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert type(build(**kw)).__name__ == cls


def test_precomputed_selects_exact():
    D = np.abs(rng.normal(size=(30, 30)))
    G = build(D, precomputed="distance")
    assert type(G).__name__ == "TraditionalGraph" and G.precomputed == "distance"


@pytest.mark.parametrize("kw,exc,msg", [
    (dict(graphtype="knn", decay=10, thresh=0),
     ValueError, "Cannot instantiate a kNNGraph with `decay=None`, `thresh=0` and `knn_max=None`. Use a TraditionalGraph instead."),
    (dict(graphtype="knn", precomputed="distance"),
     ValueError, "kNNGraph does not support precomputed values. Use `graphtype='exact'` or `precomputed=None`"),
    (dict(graphtype="knn", sample_idx=np.arange(60)),
     ValueError, "kNNGraph does not support batch correction. Use `graphtype='mnn'` or `sample_idx=None`"),
    (dict(graphtype="knn", knn=None, bandwidth=None),
     ValueError, "Either `knn` or `bandwidth` must be provided."),
    (dict(kernel_symm="invalid"),
     ValueError, "kernel_symm 'invalid' not recognized. Choose from '+', '*', 'mnn', or 'none'."),
    (dict(kernel_symm="mnn", theta=-1),
     ValueError, "theta -1 not recognized. Expected a float between 0 and 1"),
    (dict(anisotropy=2),
     ValueError, "Expected 0 <= anisotropy <= 1. Got 2"),
    (dict(graphtype="hello"),
     ValueError, "graphtype 'hello' not recognized. Choose from ['knn', 'mnn', 'exact', 'auto']"),
    (dict(graphtype="exact", decay=None),
     ValueError, "`decay` must be provided for a TraditionalGraph. For kNN kernel, use kNNGraph."),
    (dict(graphtype="exact", sample_idx=np.arange(60) % 2),
     ValueError, "TraditionalGraph does not support batch correction. Use `graphtype='mnn'` or `sample_idx=None`"),
    (dict(graphtype="mnn", sample_idx=np.arange(60) % 2, precomputed="distance"),
     ValueError, "MNNGraph does not support precomputed values. Use `graphtype='exact'` and `sample_idx=None` or `precomputed=None`"),
    (dict(graphtype="mnn", sample_idx=None),
     ValueError, "sample_idx must be given. For a graph without batch correction, use kNNGraph."),
    (dict(sample_idx=np.arange(59) % 2),
     ValueError, "sample_idx (59) must be the same length as data (60)"),
    (dict(n_landmark=60),
     ValueError, "n_landmark (60) >= n_samples (60). Use kNNGraph instead"),
    (dict(bandwidth=lambda d: 1.0, graphtype="knn"),
     NotImplementedError, "Callable bandwidth is only supported by graphtools.graphs.TraditionalGraph."),
    (dict(n_pca=-1), ValueError, "n_pca cannot be negative."),
    (dict(n_pca="hello"), ValueError, "n_pca must be an integer"),
    (dict(hello="world"), TypeError, "hello"),
    (dict(sample_idx=np.arange(60) % 2, kernel_symm="mnn", theta="a"), TypeError, "Expected `theta` as a float"),
])
def test_errors(kw, exc, msg):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(exc) as e:
            build(**kw)
    assert msg in str(e.value)


def test_precomputed_errors():
    with pytest.raises(ValueError, match="Precomputed value invalid not recognized"):
        build(np.abs(rng.normal(size=(30, 30))), precomputed="invalid")
    with pytest.raises(ValueError, match="must be a square matrix"):
        build(np.abs(rng.normal(size=(30, 20))), precomputed="distance")
    with pytest.raises(ValueError, match="should be non-negative"):
        build(-np.abs(rng.normal(size=(30, 30))), precomputed="distance")


@pytest.mark.parametrize("kw,cat,msg", [
    (dict(knn=59, decay=10), UserWarning, "Cannot set knn (59) to be greater than n_samples - 2 (58). Setting knn=58"),
    (dict(knn=10, knn_max=9, decay=10), UserWarning, "Cannot set knn_max (9) to be less than knn (10). Setting knn_max=10"),
    (dict(decay=None, bandwidth=3), UserWarning, "`bandwidth` is not used when `decay=None`."),
    (dict(kernel_symm="+", theta=0.5), UserWarning, "kernel_symm='+' but theta is not None. Setting kernel_symm='mnn'."),
    (dict(kernel_symm="mnn"), UserWarning, "kernel_symm='mnn' but theta not given. Defaulting to theta=1."),
    (dict(kernel_symm="gamma", theta=0.5), FutureWarning, "kernel_symm='gamma' is deprecated. Setting kernel_symm='mnn'"),
    (dict(gamma=0.5, kernel_symm="mnn"), FutureWarning, "gamma is deprecated. Setting theta=0.5"),
    (dict(sample_idx=np.zeros(60)), UserWarning, "Only one unique sample. Not using MNNGraph"),
    (dict(n_pca=100), RuntimeWarning, "Cannot perform PCA to 100 dimensions on data with min(n_samples, n_features) = 8"),
    (dict(n_landmark=20, n_svd=60), RuntimeWarning, "n_svd (60) >= n_samples (60) Consider using kNNGraph or lower n_svd"),
    (dict(sample_idx=np.arange(60) % 2, adaptive_k="sqrt", kernel_symm="mnn", theta=0.5), DeprecationWarning,
     "`adaptive_k` has been deprecated. Using fixed knn."),
])
def test_warnings(kw, cat, msg):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        build(**kw)
    assert any(issubclass(x.category, cat) and msg in str(x.message) for x in w), [str(x.message) for x in w]


def test_thresh_clamped_to_eps_and_params():
    G = build(decay=10, thresh=1e-30, knn_max=20)
    assert G.thresh == np.finfo(float).eps
    p = G.get_params()
    for k in ("knn", "decay", "bandwidth", "bandwidth_scale", "knn_max", "distance", "thresh", "n_jobs", "random_state",
              "verbose", "kernel_symm", "theta", "anisotropy", "n_pca"):
        assert k in p
    with pytest.raises(ValueError, match="Cannot update knn. Please create a new graph"):
        G.set_params(knn=G.knn + 1)
    with pytest.raises(ValueError, match="Cannot update decay. Please create a new graph"):
        G.set_params(decay=3)
    with pytest.raises(ValueError, match="Cannot update kernel_symm. Please create a new graph"):
        G.set_params(kernel_symm="*")
    assert G.set_params(n_jobs=4, verbose=0, random_state=3) is G and G.n_jobs == 4 and G.random_state == 3


def test_landmark_params_reset():
    G = build(n_landmark=20)
    G._clusters = np.zeros(60, dtype=int)
    G.set_params(n_landmark=25)
    assert G.n_landmark == 25 and not hasattr(G, "_clusters")


def test_pca_reduction_on_host():
    X = rng.normal(size=(80, 30))
    G = build(X, n_pca=5, random_state=1)
    assert G.data_nu.shape == (80, 5) and G.n_pca == 5
    Y = rng.normal(size=(7, 30))
    assert G.transform(Y).shape == (7, 5)
    assert G.inverse_transform(G.data_nu).shape == (80, 30)
    with pytest.raises(ValueError, match="Y must be of shape either"):
        G._check_extension_shape(rng.normal(size=(3, 9)))


def test_unsupported_metric_is_rejected_loudly():
    with pytest.raises(NotImplementedError, match="euclidean, cosine and cityblock"):
        build(distance="minkowski")
    with pytest.raises(NotImplementedError, match="euclidean, cosine and cityblock"):
        build(distance="chebyshev", graphtype="exact")


def test_mnn_to_data_not_implemented():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = build(sample_idx=np.arange(60) % 2, kernel_symm="mnn", theta=0.5)
    with pytest.raises(NotImplementedError):
        G.build_kernel_to_data(DATA)


# ------------------------------------------------------------------ switches added with the widened rows
def test_pca_front_end_without_gpu_is_the_reference_call(monkeypatch):
    """GTB_PCA=auto keeps scikit-learn's PCA when no GPU is present (host-side API use); values = sklearn's."""
    from sklearn.decomposition import PCA
    import torch
    if torch.cuda.is_available():
        pytest.skip("covers the host-only container")
    X = rng.normal(size=(80, 20))
    G = build(X, n_pca=5, random_state=3)
    ref = PCA(5, svd_solver="randomized", random_state=3).fit(X)
    assert np.array_equal(G.data_nu, ref.transform(X))
    assert np.array_equal(G.transform(X[:4]), ref.transform(X[:4]))
    monkeypatch.setenv("GTB_PCA", "bogus")
    with pytest.raises(ValueError, match="GTB_PCA"):
        build(X, n_pca=5)


def test_spectral_switch_validation(monkeypatch):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = build(n_landmark=20)
    monkeypatch.setenv("GTB_SPECTRAL", "bogus")
    with pytest.raises(ValueError, match="GTB_SPECTRAL"):
        G._spectral_impl()
    monkeypatch.setenv("GTB_SPECTRAL", "auto")
    assert G._spectral_impl() == "host"            # small input, no device kernel yet
    monkeypatch.setenv("GTB_SPECTRAL", "device")
    with pytest.raises(NotImplementedError):
        G._spectral_impl()                         # needs a built sparse kernel in HBM


def test_cosine_is_accepted_by_every_graph_type():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert build(distance="cosine").distance == "cosine"
        assert build(distance="cosine", graphtype="exact").distance == "cosine"
        assert build(distance="cosine", sample_idx=np.arange(60) % 2).distance == "cosine"
        assert build(distance="cosine", n_landmark=10, random_landmarking=True).distance == "cosine"


def test_tie_exemption_uses_the_graph_metric():
    """The comparator's k-th neighbour tie test must run in the metric of the graph (tests/parity.py)."""
    from tests.parity import _tie_exempt
    X = np.array([[1.0, 0.0], [2.0, 0.0], [0.0, 3.0], [1.0, 1.0]])
    # rows 0 and 1 point the same way: cosine distance 0 -- a tie with the self distance for k = 2
    ok_cos = _tie_exempt(np.array([0]), np.array([1]), X, 2, metric="cosine")
    ok_euc = _tie_exempt(np.array([0]), np.array([1]), X, 1, metric="euclidean")
    assert ok_cos[0] and not ok_euc[0]


def test_pickle_round_trip_and_read_pickle(tmp_path):
    """to_pickle / read_pickle (reference api.py:339-354, test_api.py:89-137) on graphs without device state."""
    import pickle
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = build(knn=4, decay=20, n_landmark=10)
    path = str(tmp_path / "graph.pkl")
    G.to_pickle(path)
    G2 = gt.read_pickle(path)
    assert type(G2).__name__ == type(G).__name__ and G2.get_params() == G.get_params()
    assert np.array_equal(G2.data, G.data)
    other = str(tmp_path / "other.pkl")
    with open(other, "wb") as f:
        pickle.dump("hello world", f)
    with pytest.warns(UserWarning, match="Returning object that is not a graphtools.base.BaseGraph"):
        assert gt.read_pickle(other) == "hello world"


def test_from_igraph_contract():
    """from_igraph (reference api.py:298-336) with a minimal stand-in for igraph.Graph.get_adjacency."""
    A = (np.abs(rng.normal(size=(12, 12))) > 1.0).astype(float)
    A = np.maximum(A, A.T)

    class _Adj:
        def __init__(self, data):
            self.data = data

    class _IG:
        def get_adjacency(self, attribute=None):
            if attribute not in (None, "weight"):
                raise ValueError("Attribute does not exist")
            return _Adj(A.tolist())

    G = gt.from_igraph(_IG(), initialize=False, verbose=0)
    assert type(G).__name__ == "TraditionalGraph" and G.precomputed == "adjacency"
    with pytest.warns(UserWarning, match="Cannot build graph from igraph with precomputed=affinity"):
        gt.from_igraph(_IG(), precomputed="affinity", initialize=False, verbose=0)
    with pytest.warns(UserWarning, match="Edge attribute nope not found. Returning unweighted graph"):
        gt.from_igraph(_IG(), attribute="nope", initialize=False, verbose=0)
