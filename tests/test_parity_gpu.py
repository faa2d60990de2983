"""Parity of the CUDA path against the golden vectors of the unmodified reference and against the
CPU oracle on seeded inputs (run on the B200 box through the public Graph API -> C ABI)."""
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


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_golden(name):
    case = Case(name)
    G = _build(case)
    assert type(G).__name__ == case.cls
    p = case.params
    thresh = p.get("thresh", 1e-4) if p.get("decay", 40) is not None else None
    K = G.kernel
    assert sparse.isspmatrix_csr(K) and K.dtype == np.float64 and K.indices.dtype == np.int32
    assert K.has_sorted_indices
    r = compare_sparse(K, case.mat("K"), thresh=thresh, what=name + ".K")
    compare_sparse(G.diff_op, case.mat("P"), thresh=None if r["n_exempt"] == 0 else thresh, what=name + ".P",
                   rtol=1e-5 if r["n_exempt"] == 0 else 1e-3)
    assert np.allclose(G.kernel_degree, case.z["degree"], rtol=1e-5 if r["n_exempt"] == 0 else 1e-3)
    # raw kernel through the public build_kernel() of a fresh, uninitialised graph
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G2 = gt.Graph(case.X, n_jobs=-1, verbose=0, initialize=False, **p)
        R = G2.build_kernel().to_scipy()
    compare_sparse(R, case.mat("R"), thresh=thresh, what=name + ".R")
    if "Y" in case.z.files:
        Kyx = G.build_kernel_to_data(case.z["Y"])
        compare_sparse(Kyx, case.mat("Kyx"), thresh=thresh, what=name + ".Kyx")
        compare_sparse(G.extend_to_data(case.z["Y"]), case.mat("ext"), thresh=thresh, what=name + ".ext")


def test_row_sums_and_symmetry_100k():
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
    assert st["rows"] == 100_000


def test_oracle_parity_20k():
    """Full-pipeline parity against the CPU oracle on a seeded 20k x 100 mixture."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(20_000, 100, n_clusters=20, intrinsic_dim=10, seed=7)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    r = compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K")
    compare_sparse(G.diff_op, P_ref, thresh=1e-4 if r["n_exempt"] else None, what="P",
                   rtol=1e-5 if r["n_exempt"] == 0 else 1e-3)


def test_oracle_parity_isotropic():
    """High intrinsic dimension: most rows go through the radius pass."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(6000, 40, n_clusters=4, intrinsic_dim=None, seed=11)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    assert pipeline.stats()["radius_rows"] > 0
    compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K")
