"""Parity of the CUDA path against the golden vectors of the unmodified reference and against the
CPU oracle on seeded inputs (run on the B200 box through the public Graph API -> C ABI)."""
import os
import warnings

import numpy as np
import pytest
from scipy import sparse

import graphtools_b200 as gt
from graphtools_b200 import pipeline, synth
from tests.golden_util import Case, names
from tests.parity import compare_dense, compare_sparse

pytestmark = pytest.mark.gpu

KNN_CASES = [n for n in names() if Case(n).cls == "kNNGraph"]


def _build(case, **extra):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return gt.Graph(case.X, n_jobs=-1, verbose=0, **dict(case.params, **extra))


@pytest.fixture(params=["tc", "tc16", "tch", "tch1", "simt"])
def impl(request, monkeypatch):
    monkeypatch.setenv("GTB_SEARCH_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_golden(name, impl):
    case = Case(name)
    G = _build(case)
    assert type(G).__name__ == case.cls
    p = case.params
    thresh = p.get("thresh", 1e-4) if p.get("decay", 40) is not None else None
    K = G.kernel
    assert sparse.isspmatrix_csr(K) and K.dtype == np.float64 and K.indices.dtype == np.int32
    assert K.has_sorted_indices
    tie = None
    if p.get("decay", 40) is None or p.get("knn_max") is not None:
        k_eff = (p["knn"] if p.get("decay", 40) is None else p["knn_max"]) + 1
        tie = dict(X=case.X, knn=k_eff, metric=p.get("distance", "euclidean"))
    r = compare_sparse(K, case.mat("K"), thresh=thresh, what=name + ".K", tie=tie)
    if r["n_exempt"] == 0:
        compare_sparse(G.diff_op, case.mat("P"), what=name + ".P")
        assert np.allclose(G.kernel_degree, case.z["degree"], rtol=1e-5)
    else:
        # a tie / threshold-boundary entry changes its row's normalisation: check P on the rows whose
        # kernel row agrees with the reference
        Pg, Pr, Kr = G.diff_op, case.mat("P"), case.mat("K")

        def row_same(i):
            a, b = K[i], Kr[i]
            return (a.nnz == b.nnz and np.array_equal(a.indices, b.indices)
                    and np.allclose(a.data, b.data, rtol=1e-5, atol=0))
        same = [i for i in range(K.shape[0]) if row_same(i)]
        assert len(same) > 0.9 * Pg.shape[0]
        compare_sparse(Pg[same], Pr[same], what=name + ".P[agreeing rows]")
    # raw kernel through the public build_kernel() of a fresh, uninitialised graph
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G2 = gt.Graph(case.X, n_jobs=-1, verbose=0, initialize=False, **p)
        R = G2.build_kernel().to_scipy()
    compare_sparse(R, case.mat("R"), thresh=thresh, what=name + ".R", tie=tie)
    if "Y" in case.z.files:
        Kyx = G.build_kernel_to_data(case.z["Y"])
        compare_sparse(Kyx, case.mat("Kyx"), thresh=thresh, what=name + ".Kyx")
        compare_sparse(G.extend_to_data(case.z["Y"]), case.mat("ext"), thresh=thresh, what=name + ".ext")


def test_row_sums_and_symmetry_100k(impl):
    """Size-independent properties at the BASELINE config-2 size (100k x 100)."""
    X, _ = synth.gaussian_mixture(100_000, 100, n_clusters=20, intrinsic_dim=10, seed=0)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    K, P = G.kernel, G.diff_op
    assert abs(K - K.T).max() == 0.0
    assert (K.diagonal() == 1.0).all()
    assert np.allclose(np.asarray(P.sum(1)).ravel(), 1.0, rtol=0, atol=1e-12)
    assert K.data.min() >= 1e-4 / 2 - 1e-12 and K.data.max() <= 1.0
    assert np.array_equal(K.indptr, P.indptr) and np.array_equal(K.indices, P.indices)
    st = pipeline.stats()
    assert st["rows"] == 100_000 and st["impl"] == impl


def test_oracle_parity_20k(impl):
    """Full-pipeline parity against the CPU oracle on a seeded 20k x 100 mixture."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(20_000, 100, n_clusters=20, intrinsic_dim=10, seed=7)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    r = compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K")
    compare_sparse(G.diff_op, P_ref, thresh=1e-4 if r["n_exempt"] else None, what="P",
                   rtol=1e-5 if r["n_exempt"] == 0 else 1e-3)


def test_oracle_parity_isotropic(impl):
    """High intrinsic dimension: most rows go through the radius pass."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(6000, 100, n_clusters=2, intrinsic_dim=None, seed=11)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    st = pipeline.stats()
    compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K")
    compare_sparse(G.diff_op, P_ref, what="P")
    assert st["radius_rows"] > 0, "stress case did not exercise the radius pass"


EXACT_CASES = [n for n in names() if Case(n).cls == "TraditionalGraph"]
MNN_CASES = [n for n in names() if Case(n).cls == "MNNGraph"]
LANDMARK_CASES = [n for n in names() if "Landmark" in Case(n).cls]


@pytest.mark.parametrize("name", EXACT_CASES)
def test_exact_golden(name):
    case = Case(name)
    G = _build(case)
    assert type(G).__name__ == case.cls
    p = case.params
    thresh = p.get("thresh", 1e-4)
    K = G.kernel
    assert isinstance(K, np.ndarray) and K.dtype == np.float64
    compare_dense(K, case.mat("K"), thresh=thresh, what=name + ".K")
    compare_dense(G.diff_op, case.mat("P"), thresh=thresh, what=name + ".P")
    assert np.allclose(G.kernel_degree, case.z["degree"], rtol=1e-5)
    compare_dense(G.build_kernel().cpu().numpy(), case.mat("R"), thresh=thresh, what=name + ".R")
    if "Y" in case.z.files:
        compare_dense(G.build_kernel_to_data(case.z["Y"]), case.mat("Kyx"), thresh=thresh, what=name + ".Kyx")
        compare_dense(G.extend_to_data(case.z["Y"]), case.mat("ext"), thresh=thresh, what=name + ".ext")


def _mnn_tie_points(X, sample_idx, knn, rel=1e-6):
    """Points whose binary MNN rows may legitimately differ from the reference: a k-th / (k+1)-th neighbour tie
    (within ``rel``, the north_star exemption) in one of the per-batch cuts of graphs.py:1868-1920 -- the in-sample
    (knn + 1, self included) cut of the point's own batch or the knn cut against another batch -- plus the two tied
    neighbours themselves (their symmetrised within-batch row sum, hence their cross-batch scale, changes too)."""
    X = np.asarray(X, dtype=np.float64)
    sample_idx = np.asarray(sample_idx)
    touched = set()
    for s in np.unique(sample_idx):
        ref = np.flatnonzero(sample_idx == s)
        for i in range(X.shape[0]):
            d = np.sqrt(((X[ref] - X[i]) ** 2).sum(1))
            k = knn + 1 if sample_idx[i] == s else knn
            if k >= len(ref):
                continue
            order = np.argsort(d, kind="stable")
            a, b = d[order[k - 1]], d[order[k]]
            if b - a <= rel * max(b, 1e-300):
                touched.update((i, int(ref[order[k - 1]]), int(ref[order[k]])))
    return touched


@pytest.mark.parametrize("name", MNN_CASES)
def test_mnn_golden(name):
    """MNN kernels against the unmodified reference, decay and binary alike: structure exact and values within
    rtol 1e-5 everywhere except on rows / columns touched by a k-th neighbour tie inside one batch block."""
    case = Case(name)
    G = _build(case)
    assert type(G).__name__ == case.cls
    p = case.params
    binary = p.get("decay", 40) is None
    thresh = None if binary else p.get("thresh", 1e-4)
    K, K_ref = G.kernel, case.mat("K")
    if isinstance(K_ref, np.ndarray):
        # thresh == 0: exact sub-graphs, dense kernel (api.py:207-209, graphs.py:1901-1902)
        assert isinstance(K, np.ndarray) and K.dtype == np.float64
        compare_dense(K, K_ref, what=name + ".K")
        compare_dense(G.diff_op, case.mat("P"), what=name + ".P")
        assert np.allclose(G.kernel_degree, case.z["degree"], rtol=1e-5)
        compare_dense(G.build_kernel().cpu().numpy(), case.mat("R"), what=name + ".R")
        return
    tied = sorted(_mnn_tie_points(case.X, p["sample_idx"], p["knn"])) if binary else []
    assert len(tied) <= 0.01 * K.shape[0], "too many tie-affected points for a meaningful check"
    if tied:
        keep = np.ones(K.shape[0], dtype=bool)
        keep[tied] = False
        D = sparse.diags(keep.astype(np.float64))
        K, K_ref = (D @ K @ D).tocsr(), (D @ K_ref @ D).tocsr()
        K.eliminate_zeros(); K_ref.eliminate_zeros()
    r = compare_sparse(K, K_ref, thresh=thresh, what=name + ".K")
    assert abs(G.kernel - G.kernel.T).max() == 0.0
    if r["n_exempt"] == 0 and not tied:
        compare_sparse(G.diff_op, case.mat("P"), what=name + ".P")
        assert np.allclose(G.kernel_degree, case.z["degree"], rtol=1e-5)
    R = G.build_kernel().to_scipy()
    R_ref = case.mat("R")
    if tied:
        R, R_ref = (D @ R @ D).tocsr(), (D @ R_ref @ D).tocsr()
        R.eliminate_zeros(); R_ref.eliminate_zeros()
    compare_sparse(R, R_ref, thresh=thresh, what=name + ".R")


def test_kernel_sanity_warnings():
    """a11 (base.py:551-554): 'K should be symmetric' for an unsymmetrised adaptive-bandwidth kernel, 'K should have a
    non-zero diagonal' for an affinity without one -- and neither for an ordinary symmetrised graph."""
    case = Case("digits_nosym")
    with pytest.warns(RuntimeWarning, match="K should be symmetric"):
        gt.Graph(case.X, n_jobs=-1, verbose=0, **case.params)
    with pytest.warns(RuntimeWarning, match="K should have a non-zero diagonal"):
        gt.Graph(np.zeros((10, 10)), precomputed="affinity", n_pca=None, verbose=0)
    A = sparse.random(300, 300, density=0.02, random_state=3, format="csr")
    A = ((A + A.T) / 2).tolil()
    A.setdiag(0)
    A = A.tocsr(); A.eliminate_zeros()
    with pytest.warns(RuntimeWarning, match="K should have a non-zero diagonal"):
        G = gt.Graph(A, precomputed="affinity", n_pca=None, verbose=0, thresh=0)
    assert abs(sparse.csr_matrix(G.kernel) - A).max() < 1e-15
    # the flags of the sparse finalising kernel themselves (csrc/sparse.cu): bit 0 asymmetric, bit 1 no diagonal
    M = sparse.csr_matrix(np.array([[1.0, 0.5, 0.0], [0.1, 1.0, 0.0], [0.0, 0.0, 0.0]]))
    _, _, _, flags = pipeline.symmetrize_normalize(pipeline.csr_from_scipy(M), None)
    assert flags & 1 and flags & 2
    _, _, _, flags = pipeline.symmetrize_normalize(pipeline.csr_from_scipy(M), "+")
    assert not (flags & 1) and flags & 2
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        gt.Graph(Case("mix_knn").X, knn=5, decay=40, verbose=0)
        gt.Graph(Case("small_exact").X, graphtype="exact", knn=5, decay=40, verbose=0)


@pytest.mark.parametrize("name", LANDMARK_CASES)
def test_landmark_golden(name):
    case = Case(name)
    p = case.params
    G = _build(case)
    assert type(G).__name__ == case.cls
    if p.get("random_landmarking"):
        # random landmarking is deterministic given the seed: clusters must match bit for bit
        assert np.array_equal(G.clusters, case.z["clusters"])
    else:
        # spectral clusters come from seeded host SVD + k-means on a K that is only rtol-close:
        # report agreement, then grade the operator with the reference's clusters injected
        agree = float(np.mean(G.clusters == case.z["clusters"]))
        print("spectral cluster agreement: %.4f" % agree)
        G.clusters = case.z["clusters"]
    op = G.landmark_op
    assert isinstance(op, np.ndarray) and op.dtype == np.float64
    compare_dense(op, case.mat("landmark_op"), what=name + ".landmark_op", rtol=1e-5)
    compare_sparse(G.transitions, case.mat("transitions"), what=name + ".transitions")
    if "Y" in case.z.files:
        compare_sparse(G.extend_to_data(case.z["Y"]), case.mat("ext"), thresh=1e-4, what=name + ".ext")


def test_diff_aff_and_interpolate():
    """diff_aff (base.py:668-698) and interpolate (base.py:1195-1229) against the oracle."""
    from oracle import graph_oracle as go
    case = Case("mix_knn")
    G = _build(case)
    K_ref = case.mat("K")
    compare_sparse(G.diff_aff, go.diff_aff(K_ref), what="diff_aff")
    Y = case.z["Y"]
    T = np.random.default_rng(0).normal(size=(K_ref.shape[0], 3))
    ref = case.mat("ext").dot(T)
    got = G.interpolate(T, Y=Y)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-12)


def test_pickle_roundtrip(tmp_path):
    import pickle
    case = Case("digits_knn5_decay40")
    G = _build(case)
    G.to_pickle(str(tmp_path / "g.pkl"))
    with open(tmp_path / "g.pkl", "rb") as f:
        G2 = pickle.load(f)
    assert (G2.kernel != G.kernel).nnz == 0 and (G2.diff_op != G.diff_op).nnz == 0


def test_pickled_landmark_graph_keeps_working(tmp_path):
    """A graph pickled BEFORE its landmark operator was read (device handles are dropped, base.py:887-902 contract) must
    rebuild its device state on demand: landmark_op, extend_to_data, interpolate and diffuse on the loaded copy."""
    import pickle
    case = Case("mix_landmark_random")
    G = _build(case)
    G.kernel
    blob = pickle.dumps(G)
    G2 = pickle.loads(blob)
    assert not any(k.startswith("_dev_") for k in G2.__dict__)
    op = G2.landmark_op
    compare_dense(op, case.mat("landmark_op"), what="landmark_op after unpickling")
    compare_sparse(G2.transitions, case.mat("transitions"), what="transitions after unpickling")
    Y = case.z["Y"]
    compare_sparse(G2.extend_to_data(Y), case.mat("ext"), thresh=1e-4, what="ext after unpickling")
    x = np.random.default_rng(0).normal(size=(G2.kernel.shape[0], 2))
    assert np.allclose(G2.diffuse(x, t=2), G.diff_op.dot(G.diff_op.dot(x)), rtol=1e-12, atol=1e-14)
    sig = np.random.default_rng(1).normal(size=(op.shape[0], 2))
    assert np.allclose(G2.interpolate(sig), G2.transitions.dot(sig), rtol=1e-12, atol=1e-14)
    # a graph pickled after everything was read round-trips bit for bit
    G3 = pickle.loads(pickle.dumps(G2))
    assert np.array_equal(G3.landmark_op, op) and (G3.kernel != G2.kernel).nnz == 0


def test_duplicate_points_warn_and_match(impl):
    """Exact duplicates: zero distances, bandwidth floor, RuntimeWarning (reference graphs.py:787-817,
    test/test_knn.py:53-68)."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(1500, 30, n_clusters=3, intrinsic_dim=6, seed=21)
    X = np.vstack([X, X[:9]])
    with pytest.warns(RuntimeWarning, match=r"Detected zero distance between samples ([0-9and,\s]*). Consider "
                                            r"removing duplicates"):
        G = gt.Graph(X, knn=5, decay=10, thresh=1e-4, verbose=0)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=10, thresh=1e-4)
    compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K")
    compare_sparse(G.diff_op, P_ref, what="P")
    X2 = np.vstack([X, X[:30]])
    with pytest.warns(RuntimeWarning, match=r"Detected zero distance between ([0-9]*) pairs of samples"):
        gt.Graph(X2, knn=5, decay=10, thresh=1e-4, verbose=0)


def test_bit_identical_across_search_implementations():
    """The certified select-then-evaluate pipeline makes the result independent of the fast pass."""
    import os
    X, _ = synth.gaussian_mixture(30_000, 64, n_clusters=10, intrinsic_dim=8, seed=5)
    out = {}
    for impl in ("tc", "tc16", "tch", "tch1", "simt"):
        os.environ["GTB_SEARCH_IMPL"] = impl
        try:
            G = gt.Graph(X, knn=5, decay=40, verbose=0)
            out[impl] = (G.kernel, G.diff_op)
        finally:
            os.environ.pop("GTB_SEARCH_IMPL", None)
    for impl in ("tc16", "tch", "tch1", "simt"):
        assert (out[impl][0] != out["tc"][0]).nnz == 0
        assert (out[impl][1] != out["tc"][1]).nnz == 0


# ------------------------------------------------------------------ wider coverage of the class surface
def test_per_row_bandwidth_and_knn_max_on_isotropic_data(impl):
    """Vector bandwidth (graphs.py:895-897) and a knn_max cut that bites on rows going through the radius
    pass (graphs.py:950-962)."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(4000, 60, n_clusters=2, intrinsic_dim=None, seed=13)
    bw = np.random.default_rng(1).uniform(9.5, 11.0, size=4000)
    g = go.KnnOracle(X.astype(np.float64), knn=5, decay=20, thresh=1e-3, bandwidth=bw)
    K_ref = go.finish_kernel(g.kernel(), "+")
    G = gt.Graph(X, knn=5, decay=20, thresh=1e-3, bandwidth=bw, verbose=0)
    compare_sparse(G.kernel, K_ref, thresh=1e-3, what="vector bandwidth")
    K_ref2, _ = go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4, knn_max=15)
    G2 = gt.Graph(X, knn=5, decay=40, thresh=1e-4, knn_max=15, verbose=0)
    compare_sparse(G2.kernel, K_ref2, thresh=1e-4, what="knn_max", tie=dict(X=X, knn=16))


@pytest.mark.parametrize("precomputed", ["distance", "affinity", "adjacency"])
def test_precomputed_exact_graphs(precomputed):
    """graphs.py:1532-1549: element-wise kernels of user-supplied matrices."""
    from scipy.spatial.distance import pdist, squareform
    X, _ = synth.gaussian_mixture(300, 10, n_clusters=3, intrinsic_dim=4, seed=2)
    D = squareform(pdist(X.astype(np.float64)))
    if precomputed == "distance":
        data = D
        bw = np.max(np.partition(D, 6, axis=1)[:, :6], axis=1)
        K = np.exp(-1 * np.power((D.T / bw).T, 40)); K[K < 1e-4] = 0
    elif precomputed == "affinity":
        data = np.exp(-D / D.mean()); K = data.copy(); K[K < 1e-4] = 0
    else:
        data = (D < np.percentile(D, 5)).astype(float); np.fill_diagonal(data, 0)
        K = data.copy(); np.fill_diagonal(K, 1)
    K = (K + K.T) / 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(data, precomputed=precomputed, knn=5, decay=40 if precomputed == "distance" else None,
                     verbose=0)
    assert type(G).__name__ == "TraditionalGraph"
    compare_dense(G.kernel, K, what=precomputed)
    P = K / K.sum(1)[:, None]
    compare_dense(G.diff_op, P, what=precomputed + ".P")
    with pytest.raises(ValueError, match="Cannot extend kernel on precomputed graph"):
        G.build_kernel_to_data(X)


def test_exact_landmark_and_mnn_landmark_graphs():
    """Composite classes (graphs.py:1969-1979): TraditionalLandmarkGraph and MNNLandmarkGraph."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(900, 20, n_clusters=4, intrinsic_dim=5, seed=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, graphtype="exact", knn=5, decay=20, thresh=1e-3, n_landmark=60, random_landmarking=True,
                     random_state=1, verbose=0)
    assert type(G).__name__ == "TraditionalLandmarkGraph"
    K_ref = go.finish_kernel(go.exact_kernel(X.astype(np.float64), knn=5, decay=20, thresh=1e-3))
    clusters = go.random_landmark_clusters(X.astype(np.float64), 60, 1)
    assert np.array_equal(G.clusters, clusters)
    op, pnm = go.landmark_operator(K_ref, clusters)
    compare_dense(G.landmark_op, op, what="exact landmark_op")
    compare_dense(np.asarray(G.transitions), np.asarray(pnm), what="exact transitions")
    idx = np.arange(900) % 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G2 = gt.Graph(X, sample_idx=idx, kernel_symm="mnn", theta=0.6, knn=5, decay=20, thresh=1e-3, n_landmark=50,
                      random_landmarking=True, random_state=2, verbose=0)
    assert type(G2).__name__ == "MNNLandmarkGraph"
    K2 = go.finish_kernel(go.mnn_kernel(X.astype(np.float64), idx, knn=5, decay=20, thresh=1e-3), "mnn", 0.6)
    compare_sparse(G2.kernel, K2, thresh=1e-3, what="mnn K")
    op2, pnm2 = go.landmark_operator(K2, go.random_landmark_clusters(X.astype(np.float64), 50, 2))
    compare_dense(G2.landmark_op, op2, what="mnn landmark_op")
    compare_sparse(G2.transitions, pnm2, what="mnn transitions")


@pytest.mark.parametrize("knn,impl_used", [(40, "tch1"), (70, "simt")])
def test_large_knn(knn, impl_used):
    """knn up to 55 fits the one-product flavour's candidate list of 64 (tensor cores); beyond that the build falls
    back to the CUDA-core kernel (S = 128)."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(3000, 30, n_clusters=3, intrinsic_dim=6, seed=8)
    K_ref, _ = go.knn_graph(X.astype(np.float64), knn=knn, decay=None)
    G = gt.Graph(X, knn=knn, decay=None, verbose=0)
    assert pipeline.stats()["impl"] == impl_used
    compare_sparse(G.kernel, K_ref, what="knn=%d binary" % knn, tie=dict(X=X, knn=knn + 1))


def test_float64_inputs_are_evaluated_in_float64(impl):
    """Inputs that are NOT float32-exact (here: PCA output, the usual PHATE entry n_pca=...): the fast pass
    runs on a float32 copy centred in float64, every value that reaches the output is computed from the
    float64 rows, so parity with the float64 reference path holds at rtol 1e-5."""
    from oracle import graph_oracle as go
    rng = np.random.default_rng(3)
    X, _ = synth.gaussian_mixture(4000, 300, n_clusters=6, intrinsic_dim=12, seed=31)
    X = X.astype(np.float64) * (1 + 1e-9 * rng.normal(size=X.shape))      # not representable in float32
    G = gt.Graph(X, n_pca=30, random_state=5, knn=5, decay=40, thresh=1e-4, verbose=0)
    assert G.data_nu.dtype == np.float64 and G.data_nu.shape == (4000, 30)
    K_ref, P_ref = go.knn_graph(G.data_nu, knn=5, decay=40, thresh=1e-4)
    r = compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K (float64 input)")
    assert r["max_rel"] < 1e-8
    compare_sparse(G.diff_op, P_ref, what="P (float64 input)")
    Y = G.data_nu[:50] + 1e-3
    Kyx = go.KnnOracle(G.data_nu, knn=5, decay=40, thresh=1e-4).kernel_to_data(Y)
    compare_sparse(G.build_kernel_to_data(Y), Kyx, thresh=1e-4, what="Kyx (float64 input)")


@pytest.mark.parametrize("n,d,knn", [(12, 3, 3), (129, 1, 5), (257, 2, 10), (40, 7, 38), (1000, 103, 5), (600, 120, 4),
                                     (500, 200, 5)])
def test_edge_shapes(n, d, knn, impl):
    """Tiny / ragged shapes: fewer points than a tile, one feature, knn at its n-2 ceiling, feature counts
    at and beyond the tensor-core operand limits (falls back to the other kernel flavours)."""
    from oracle import graph_oracle as go
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, knn=knn, decay=15, thresh=1e-3, verbose=0)
        K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=knn, decay=15, thresh=1e-3)
    compare_sparse(G.kernel, K_ref, thresh=1e-3, what="K %s" % ((n, d, knn),))
    compare_sparse(G.diff_op, P_ref, what="P")
    y = X[:1] + np.float32(0.1)
    Kyx = go.KnnOracle(X.astype(np.float64), knn=knn, decay=15, thresh=1e-3).kernel_to_data(y.astype(np.float64))
    compare_sparse(G.build_kernel_to_data(y), Kyx, thresh=1e-3, what="single query row")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Gb = gt.Graph(X, knn=min(knn, 5), decay=None, verbose=0)
        Kb, _ = go.knn_graph(X.astype(np.float64), knn=min(knn, 5), decay=None)
    compare_sparse(Gb.kernel, Kb, what="binary", tie=dict(X=X, knn=min(knn, 5) + 1))


def test_host_result_pool_recycles_without_aliasing():
    """K / P live in recycled page-locked blocks (hostpool.py): a graph that is still referenced keeps its memory,
    a dropped one hands its block to the next build, and the values are the same either way."""
    import gc
    from graphtools_b200 import hostpool
    X, _ = synth.gaussian_mixture(120_000, 40, n_clusters=8, intrinsic_dim=8, seed=12)
    X2 = X[::-1].copy()
    G1 = gt.Graph(X, knn=5, decay=40, verbose=0)
    K1, P1 = G1.kernel, G1.diff_op
    assert any(not b.free for b in hostpool._local), "the result did not come from the pool"
    k1_copy, p1_copy = K1.data.copy(), P1.data.copy()
    G2 = gt.Graph(X2, knn=5, decay=40, verbose=0)               # K1 is alive: must not be overwritten
    K2 = G2.kernel
    assert np.array_equal(K1.data, k1_copy) and np.array_equal(P1.data, p1_copy)
    assert not np.shares_memory(K1.data, K2.data)
    n_blocks = len(hostpool._local)
    k2_copy = K2.data.copy()
    del G2, K2
    gc.collect()
    G3 = gt.Graph(X2, knn=5, decay=40, verbose=0)               # reuses G2's blocks
    K3 = G3.kernel
    assert len(hostpool._local) == n_blocks
    assert np.array_equal(K3.data, k2_copy)
    assert np.array_equal(K1.data, k1_copy)
    os.environ["GTB_HOST_POOL"] = "0"
    try:
        K4 = gt.Graph(X2, knn=5, decay=40, verbose=0).kernel
    finally:
        os.environ.pop("GTB_HOST_POOL")
    assert np.array_equal(K4.data, k2_copy) and np.array_equal(K4.indices, K3.indices)


@pytest.mark.parametrize("d,iso", [(300, False), (200, True), (510, False)])
def test_wide_data_runs_on_tensor_cores(d, iso):
    """d beyond the resident query tile of the one-product flavour (d + 2 > 128): the fp16x2 flavour streams the
    operand rows in chunks; kNN graph against the oracle, incl. an isotropic case that takes the radius pass."""
    from oracle import graph_oracle as go
    n = 4000
    X, _ = synth.gaussian_mixture(n, d, n_clusters=4, intrinsic_dim=None if iso else 10, seed=17)
    decay = 10 if iso else 40
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=decay, thresh=1e-4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, knn=5, decay=decay, thresh=1e-4, verbose=0)
        K = G.kernel
    st = pipeline.stats()
    assert st["impl"] == "tch", st
    if iso:
        assert st["radius_rows"] > 0
    r = compare_sparse(K, K_ref, thresh=1e-4, what="K (d=%d)" % d)
    if r["n_exempt"] == 0:
        compare_sparse(G.diff_op, P_ref, what="P (d=%d)" % d)
