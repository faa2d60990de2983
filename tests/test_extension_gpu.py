"""GPU parity of the "next" rows of SURVEY section 8f: device-resident out-of-sample extension, interpolation and
diffusion (sparse x dense product, csrc/spmm.cu).  The oracle for these is scipy's own CSR product on the
matrices the graph returns -- the arithmetic the reference performs at base.py:1229."""
import warnings

import numpy as np
import pytest
from scipy import sparse

import graphtools_b200 as gt
from graphtools_b200 import pipeline, synth

pytestmark = pytest.mark.gpu


def _rand_csr(n, m, mean_nnz, seed):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        k = 0 if i % 17 == 0 else int(rng.integers(1, 2 * mean_nnz))
        rows.append(np.sort(rng.choice(m, size=min(k, m), replace=False)))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    indices = np.concatenate(rows).astype(np.int32) if indptr[-1] else np.zeros(0, np.int32)
    data = rng.standard_normal(indptr[-1])
    return sparse.csr_matrix((data, indices, indptr), shape=(n, m))


@pytest.mark.parametrize("f", [1, 2, 7, 64, 65, 100, 110, 128, 129, 257])
def test_spmm_bit_identical_to_scipy(f):
    """Stored-order accumulation with separate multiply and add = scipy csr_matvecs, bit for bit."""
    import torch
    A = _rand_csr(3001, 2500, 9, seed=f)
    B = np.random.default_rng(100 + f).standard_normal((2500, f))
    out = pipeline.spmm(pipeline.csr_from_scipy(A), torch.from_numpy(B).cuda()).cpu().numpy()
    ref = A.dot(B)
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())


def test_spmm_strided_operand_and_value_override():
    import torch
    A = _rand_csr(1000, 800, 6, seed=5)
    Bfull = np.random.default_rng(6).standard_normal((800, 96))
    Bd = torch.from_numpy(Bfull).cuda()[:, :50]            # row stride 96, 50 columns used
    vals = np.abs(A.data) + 1.0
    Ad = pipeline.csr_from_scipy(A)
    out = pipeline.spmm(Ad, Bd, torch.from_numpy(vals).cuda()).cpu().numpy()
    A2 = sparse.csr_matrix((vals, A.indices, A.indptr), shape=A.shape)
    assert np.array_equal(out, A2.dot(Bfull[:, :50]))


def _graph(n=6000, d=40, **kw):
    X, _ = synth.gaussian_mixture(n, d, n_clusters=6, intrinsic_dim=8, seed=11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return X, gt.Graph(X, knn=5, decay=40, verbose=0, **kw)


def test_interpolate_device_equals_host_product():
    X, G = _graph()
    Y = X[::7] + 0.01 * np.random.default_rng(0).standard_normal(X[::7].shape).astype(np.float32)
    transform = np.random.default_rng(1).standard_normal((X.shape[0], 12))
    T = G.extend_to_data(Y)
    assert sparse.isspmatrix_csr(T) and T.shape == (Y.shape[0], X.shape[0])
    assert np.allclose(np.asarray(T.sum(axis=1)).ravel(), 1.0, rtol=0, atol=1e-12)
    got = G.interpolate(transform, Y=Y)
    assert np.array_equal(got, T.dot(transform))
    # caller-supplied transitions take the same device product
    assert np.array_equal(G.interpolate(transform, transitions=T), T.dot(transform))
    # 1-D signals keep their shape
    v = transform[:, 0]
    assert np.array_equal(G.interpolate(v, transitions=T), T.dot(v))
    with pytest.raises(ValueError):
        G.interpolate(transform)


def test_diffuse_equals_repeated_host_products():
    X, G = _graph()
    sig = np.random.default_rng(2).standard_normal((X.shape[0], 20))
    P = G.diff_op
    ref = sig
    for _ in range(3):
        ref = P.dot(ref)
    assert np.array_equal(G.diffuse(sig, t=3), ref)
    assert np.array_equal(G.diffuse(sig[:, 0], t=1), P.dot(sig[:, 0]))


def test_landmark_interpolate_and_extend_device():
    X, G = _graph(n=5000, n_landmark=200, random_landmarking=True, random_state=3)
    L = G.landmark_op.shape[0]
    emb = np.random.default_rng(4).standard_normal((L, 5))
    T = G.transitions
    assert np.array_equal(G.interpolate(emb), T.dot(emb))
    Y = X[::11]
    Ty = G.extend_to_data(Y)
    assert Ty.shape == (Y.shape[0], L)
    assert np.array_equal(G.interpolate(emb, Y=Y), Ty.dot(emb))
    assert np.allclose(np.asarray(Ty.sum(axis=1)).ravel(), 1.0, rtol=0, atol=1e-12)


def test_exact_graph_interpolate_dense():
    X, _ = synth.gaussian_mixture(1500, 20, n_clusters=4, intrinsic_dim=6, seed=9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, graphtype="exact", knn=5, decay=40, verbose=0)
    Y = X[:100]
    T = G.extend_to_data(Y)
    tr = np.random.default_rng(5).standard_normal((X.shape[0], 3))
    assert np.allclose(G.interpolate(tr, Y=Y), T.dot(tr), rtol=1e-12, atol=1e-14)
    sig = np.random.default_rng(6).standard_normal((X.shape[0], 4))
    assert np.allclose(G.diffuse(sig, t=2), G.diff_op.dot(G.diff_op.dot(sig)), rtol=1e-11, atol=1e-13)
