"""PCA front-end on the device (reference graphtools/base.py:227-294).

The reference reduces the input with scikit-learn before any graph is built:
``PCA(n_pca, svd_solver="randomized", random_state)`` for dense data, ``TruncatedSVD(n_pca, random_state)`` for
sparse data, then ``data_nu = data_pca.transform(data)``.  Both are sklearn's randomized SVD (Halko et al.).  This
module runs the same algorithm in HBM and in float64:

* the Gaussian test matrix is drawn on the HOST from the same ``numpy.random.RandomState`` stream sklearn uses
  (it is only [n_features, n_pca + 10]), so the computation starts from the same point as the reference's;
* dense data: products are library GEMMs (plumbing); sparse data: ``X @ Q`` and ``X.T @ Q`` are the CSR x dense
  kernel (csrc/spmm.cu) on X and on the CSR of its transpose -- the accumulation order of scipy's csr / csc products;
* the power-iteration normaliser is Cholesky-QR instead of sklearn's LU: any normaliser spans the same subspace,
  so the final factors agree with sklearn's to rounding (checked in tests/test_pca_gpu.py to 1e-7 relative);
* sign convention ``svd_flip(u_based_decision=False)`` and every fitted attribute (``components_``, ``mean_``,
  ``explained_variance_[_ratio_]``, ``singular_values_``, ``noise_variance_``) are filled into a genuine sklearn
  estimator object, so ``transform`` / ``inverse_transform`` of out-of-sample points keep working on the host.

Selected by ``GTB_PCA`` = ``auto`` (default: device whenever a GPU is present) | ``device`` | ``host`` (the reference's
own sklearn calls).
"""
import numpy as np
import torch
from scipy import sparse

from . import pipeline
from .spectral import _cholesky_qr


def _check_random_state(seed):
    from sklearn.utils import check_random_state
    return check_random_state(seed)


def _randomized_svd(matvec, rmatvec, shape, n_components, n_oversamples, n_iter, random_state):
    """sklearn.utils.extmath._randomized_svd (flip_sign=False) on operators: matvec(Q) = M @ Q [m, k],
    rmatvec(Q) = M.T @ Q [n, k]; returns device tensors U [m, c], s [c], Vt [c, n]."""
    m, n = shape
    n_random = n_components + n_oversamples
    if n_iter == "auto":
        n_iter = 7 if n_components < 0.1 * min(shape) else 4
    transpose = m < n
    if transpose:
        matvec, rmatvec, m, n = rmatvec, matvec, n, m
    Q = torch.from_numpy(random_state.normal(size=(n, n_random))).to(pipeline._dev())
    normalise = n_iter > 2
    for _ in range(int(n_iter)):
        Q = matvec(Q)
        if normalise:
            Q = _cholesky_qr(Q, passes=1)
        Q = rmatvec(Q)
        if normalise:
            Q = _cholesky_qr(Q, passes=1)
    Q = _cholesky_qr(matvec(Q), passes=2)                 # orthonormal basis of the range, [m, k]
    B = rmatvec(Q).T                                      # Q.T @ M, [k, n]
    Uhat, s, Vt = torch.linalg.svd(B, full_matrices=False)
    U = Q @ Uhat
    if transpose:
        U, Vt = Vt.T, U.T
    return U[:, :n_components], s[:n_components], Vt[:n_components]


def _svd_flip_v(U, Vt):
    """svd_flip(u_based_decision=False): the largest-magnitude entry of every row of Vt is positive."""
    piv = Vt.abs().argmax(dim=1)
    signs = torch.sign(Vt[torch.arange(Vt.shape[0], device=Vt.device), piv])
    signs[signs == 0] = 1.0
    return U * signs[None, :], Vt * signs[:, None]


def fit_transform_dense(X, n_components, random_state):
    """(fitted sklearn PCA object, data_nu as float64 CUDA tensor) for dense ``X`` (numpy / tensor)."""
    from sklearn.decomposition import PCA
    Xd = pipeline.to_device(np.asarray(X, dtype=np.float64) if not isinstance(X, torch.Tensor) else X.double())
    n, d = Xd.shape
    mean = Xd.mean(dim=0)
    Xc = Xd - mean
    rs = _check_random_state(random_state)
    U, S, Vt = _randomized_svd(lambda Q: Xc @ Q, lambda Q: Xc.T @ Q, (n, d), n_components, 10, "auto", rs)
    U, Vt = _svd_flip_v(U, Vt)
    op = PCA(n_components, svd_solver="randomized", random_state=random_state)
    total_var = float((Xc * Xc).sum().item()) / (n - 1)
    ev = (S ** 2 / (n - 1)).cpu().numpy()
    op.mean_ = mean.cpu().numpy()
    op.components_ = Vt.cpu().numpy()
    op.n_components_ = int(n_components)
    op.n_samples_ = int(n)
    op.n_features_in_ = int(d)
    op.explained_variance_ = ev
    op.explained_variance_ratio_ = ev / total_var
    op.singular_values_ = S.cpu().numpy().copy()
    if n_components < min(n, d):
        op.noise_variance_ = (total_var - ev.sum()) / (min(n, d) - n_components)
    else:
        op.noise_variance_ = 0.0
    op._fit_svd_solver = "randomized"
    # _BasePCA._transform: X @ components_.T - mean_ @ components_.T
    data_nu = Xd @ Vt.T - (mean[None, :] @ Vt.T)
    return op, data_nu


def fit_transform_sparse(X, n_components, random_state):
    """(fitted sklearn TruncatedSVD object, data_nu as float64 CUDA tensor) for scipy-sparse ``X``."""
    from sklearn.decomposition import TruncatedSVD
    from sklearn.utils.sparsefuncs import mean_variance_axis
    X = sparse.csr_matrix(X, dtype=np.float64)
    n, d = X.shape
    if n_components > d:
        raise ValueError("n_components({}) must be <= n_features({}).".format(n_components, d))
    A = pipeline.csr_from_scipy(X)
    At = pipeline.csr_from_scipy(sparse.csr_matrix(X.T))      # CSR of the transpose = scipy's csc product order
    rs = _check_random_state(random_state)
    U, S, Vt = _randomized_svd(lambda Q: pipeline.spmm(A, Q), lambda Q: pipeline.spmm(At, Q), (n, d), n_components,
                               10, 5, rs)
    U, Vt = _svd_flip_v(U, Vt)
    data_nu = pipeline.spmm(A, Vt.T.contiguous())              # safe_sparse_dot(X, components_.T)
    op = TruncatedSVD(n_components, random_state=random_state)
    op.components_ = Vt.cpu().numpy()
    op.n_features_in_ = int(d)
    exp_var = data_nu.var(dim=0, unbiased=False).cpu().numpy()
    _, full_var = mean_variance_axis(X, axis=0)
    op.explained_variance_ = exp_var
    op.explained_variance_ratio_ = exp_var / full_var.sum()
    op.singular_values_ = S.cpu().numpy().copy()
    return op, data_nu
