"""Device PCA front-end (graphtools_b200/pca.py; reference base.py:227-294) against scikit-learn's PCA /
TruncatedSVD with the same seed: the Gaussian test matrix comes from the same numpy stream, so the factors must
agree to rounding (float64), and a graph built on the device-reduced data must equal the oracle's graph on
sklearn-reduced data at the parity tolerances."""
import warnings

import numpy as np
import pytest
from scipy import sparse

import graphtools_b200 as gt
from graphtools_b200 import pca, synth
from oracle import graph_oracle as go
from tests.parity import compare_sparse

pytestmark = pytest.mark.gpu


def _dense(n=3000, d=300, seed=0):
    X, _ = synth.gaussian_mixture(n, d, n_clusters=6, intrinsic_dim=12, seed=seed)
    return X.astype(np.float64)


@pytest.mark.parametrize("shape", [(3000, 300), (200, 900)])
def test_dense_pca_matches_sklearn(shape):
    from sklearn.decomposition import PCA
    X = _dense(*shape)
    ref = PCA(40, svd_solver="randomized", random_state=42).fit(X)
    Z_ref = ref.transform(X)
    op, Z = pca.fit_transform_dense(X, 40, 42)
    Z = Z.cpu().numpy()
    scale = np.abs(Z_ref).max()
    assert np.allclose(op.singular_values_, ref.singular_values_, rtol=1e-9)
    assert np.allclose(op.components_, ref.components_, rtol=0, atol=1e-8)
    assert np.allclose(Z, Z_ref, rtol=0, atol=1e-8 * scale)
    assert np.allclose(op.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-9)
    assert np.isclose(op.noise_variance_, ref.noise_variance_, rtol=1e-9)
    assert np.allclose(op.mean_, ref.mean_, rtol=1e-12, atol=1e-14)
    # the filled-in estimator is a working sklearn object (host-side transform of new points)
    Y = X[:17] + 0.01
    assert np.allclose(op.transform(Y), ref.transform(Y), rtol=0, atol=1e-8 * scale)
    assert np.allclose(op.inverse_transform(op.transform(Y)), ref.inverse_transform(ref.transform(Y)), atol=1e-7 * scale)


def test_sparse_truncated_svd_matches_sklearn():
    from sklearn.decomposition import TruncatedSVD
    rng = np.random.default_rng(3)
    X = sparse.random(2500, 1200, density=0.03, random_state=5, format="csr", data_rvs=lambda k: rng.gamma(2.0, size=k))
    ref = TruncatedSVD(30, random_state=7)
    Z_ref = ref.fit_transform(X)
    op, Z = pca.fit_transform_sparse(X, 30, 7)
    Z = Z.cpu().numpy()
    scale = np.abs(Z_ref).max()
    assert np.allclose(op.singular_values_, ref.singular_values_, rtol=1e-9)
    assert np.allclose(op.components_, ref.components_, rtol=0, atol=1e-8)
    assert np.allclose(Z, Z_ref, rtol=0, atol=1e-8 * scale)
    assert np.allclose(op.explained_variance_, ref.explained_variance_, rtol=1e-8)
    assert np.allclose(op.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-8)
    assert np.allclose(op.transform(X[:9]), ref.transform(X[:9]), rtol=0, atol=1e-8 * scale)


def test_graph_with_n_pca_matches_oracle_on_sklearn_reduced_data(monkeypatch):
    from sklearn.decomposition import PCA
    monkeypatch.setenv("GTB_PCA", "device")
    X = _dense(4000, 200, seed=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, n_pca=30, knn=5, decay=40, random_state=1, verbose=0)
    Z_ref = PCA(30, svd_solver="randomized", random_state=1).fit(X).transform(X)
    assert G.data_nu.shape == (4000, 30) and np.allclose(G.data_nu, Z_ref, rtol=0, atol=1e-8 * np.abs(Z_ref).max())
    K_ref, P_ref = go.knn_graph(Z_ref, knn=5, decay=40)
    r = compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K on PCA-reduced data")
    if r["n_exempt"] == 0:
        compare_sparse(G.diff_op, P_ref, what="P on PCA-reduced data")
    # out-of-sample points given in the ambient space go through the fitted estimator
    Y = X[:50] + 0.01
    T = G.extend_to_data(Y)
    assert T.shape == (50, 4000) and np.allclose(np.asarray(T.sum(1)).ravel(), 1.0, atol=1e-12)
    # host path on request
    monkeypatch.setenv("GTB_PCA", "host")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Gh = gt.Graph(X, n_pca=30, knn=5, decay=40, random_state=1, verbose=0)
    assert np.array_equal(Gh.data_nu, Z_ref)


def test_graph_on_sparse_input_with_n_pca():
    """Sparse data goes through TruncatedSVD (base.py:251-256): device path = sklearn's values, graph builds."""
    from sklearn.decomposition import TruncatedSVD
    rng = np.random.default_rng(8)
    X = sparse.random(1500, 400, density=0.05, random_state=2, format="csr", data_rvs=lambda k: rng.gamma(2.0, size=k))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, n_pca=20, knn=5, decay=40, random_state=5, verbose=0)
    Z_ref = TruncatedSVD(20, random_state=5).fit_transform(X)
    assert np.allclose(G.data_nu, Z_ref, rtol=0, atol=1e-8 * np.abs(Z_ref).max())
    K_ref, _ = go.knn_graph(Z_ref, knn=5, decay=40)
    compare_sparse(G.kernel, K_ref, thresh=1e-4, what="K on TruncatedSVD-reduced data")
    # new points in the ambient (sparse) space are reduced by the fitted estimator on the host
    T = G.extend_to_data(X[:20])
    assert T.shape == (20, 1500)
