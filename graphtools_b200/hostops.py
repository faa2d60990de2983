"""GPU versions of the public ``symmetrize_kernel`` / ``apply_anisotropy`` methods for matrices that
live on the host (scipy CSR or ndarray) -- reference graphtools/base.py:557-592.  The graph build itself
never goes through here (it stays in HBM); these exist so the public methods keep working."""
import numpy as np
import torch
from scipy import sparse

from . import pipeline


def _upload_csr(K):
    K = sparse.csr_matrix(K)
    K.sum_duplicates()
    K.sort_indices()
    dev = pipeline._dev()
    return pipeline.DeviceCSR(torch.from_numpy(K.indptr.astype(np.int64)).to(dev),
                              torch.from_numpy(K.indices.astype(np.int32)).to(dev),
                              torch.from_numpy(K.data.astype(np.float64)).to(dev), K.shape)


def symmetrize_host_matrix(K, kernel_symm, theta):
    if kernel_symm is None:
        return K
    if sparse.issparse(K):
        Ks, _, _, _ = pipeline.symmetrize_normalize(_upload_csr(K), kernel_symm, theta, 0.0, want_p=False)
        return Ks.to_scipy()
    from .dense import symmetrize_dense
    return symmetrize_dense(torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64)).to(pipeline._dev()),
                            kernel_symm, theta).cpu().numpy()


def anisotropy_host_matrix(K, anisotropy):
    if sparse.issparse(K):
        Kd = _upload_csr(K)
        Ks, _, _, _ = pipeline.symmetrize_normalize(Kd, None, None, float(anisotropy), want_p=False)
        return Ks.to_scipy()
    from .dense import anisotropy_dense
    Kd = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64)).to(pipeline._dev())
    return anisotropy_dense(Kd, float(anisotropy)).cpu().numpy()
