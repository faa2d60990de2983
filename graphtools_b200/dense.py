"""Dense (exact-graph) device operations: thin wrappers over csrc/dense.cu."""
import numpy as np
import torch

from . import _engine as E
from . import pipeline

SYM = {"+": 0, "*": 1, "mnn": 2, None: 3}
METRIC = {"euclidean": 0, "cosine": 1, "cityblock": 2, "manhattan": 2, "l1": 2}


def _same_dtype(Xq, Xr):
    """Both operands in the precision of the data as given: float64 when either is (scipy's pdist / cdist work on
    the float64 values the reference holds), else float32."""
    is64 = Xq.dtype == torch.float64 or Xr.dtype == torch.float64
    dt = torch.float64 if is64 else torch.float32
    return Xq.to(dt).contiguous(), Xr.to(dt).contiguous(), int(is64)


def dense_affinity(Xq, Xr, bw_q, bw_r, decay, thresh, symm=None, theta=None, want_rowsum=True, metric="euclidean"):
    """[nq, nr] float64 thresholded alpha-decay affinities; symmetrised in the same sweep when
    ``bw_r`` is given (square problems)."""
    Xq, Xr, is64 = _same_dtype(Xq, Xr)
    nq, d = Xq.shape
    nr = Xr.shape[0]
    out = pipeline._empty((nq, nr), torch.float64)
    rowsum = pipeline._empty((nq,), torch.float64) if want_rowsum else None
    what = 2 if bw_r is not None else 1
    E.call("gtb_dense_kernel", Xq, nq, Xr, nr, d, is64, what, METRIC[metric], bw_q, bw_r, float(decay), float(thresh),
           SYM[symm], 0.0 if theta is None else float(theta), out, rowsum)
    return out, rowsum


def dense_distances(Xq, Xr, metric="euclidean"):
    Xq, Xr, is64 = _same_dtype(Xq, Xr)
    nq, d = Xq.shape
    nr = Xr.shape[0]
    out = pipeline._empty((nq, nr), torch.float64)
    E.call("gtb_dense_kernel", Xq, nq, Xr, nr, d, is64, 0, METRIC[metric], None, None, 0.0, 0.0, 3, 0.0, out, None)
    return out


def rowsum_dense(K):
    s = pipeline._empty((K.shape[0],), torch.float64)
    E.call("gtb_dense_rowsum", K, K.shape[0], K.shape[1], s)
    return s


def row_normalize_dense(K, rowsum=None):
    if rowsum is None:
        rowsum = rowsum_dense(K)
    out = pipeline._empty(tuple(K.shape), torch.float64)
    E.call("gtb_dense_row_scale", K, rowsum, K.shape[0], K.shape[1], out)
    return out


def anisotropy_dense(K, alpha, deg=None):
    if deg is None:
        deg = rowsum_dense(K)
    newsum = pipeline._empty((K.shape[0],), torch.float64)
    E.call("gtb_dense_anisotropy", K, deg, float(alpha), K.shape[0], newsum)
    return K


def symmetrize_dense(K, kernel_symm, theta):
    """Dense symmetrisation of an arbitrary square matrix already in HBM (host-callable helper).
    The exact-graph build does not use this: it fuses the rule into the distance tile."""
    if kernel_symm == "+":
        return (K + K.T) / 2
    if kernel_symm == "*":
        return K * K.T
    if kernel_symm == "mnn":
        return theta * torch.minimum(K, K.T) + (1 - theta) * torch.maximum(K, K.T)
    return K


def _dense_to_csr(K):
    """Dense [n, m] device matrix -> DeviceCSR of its non-zeros (used to feed the landmark kernels)."""
    nz = K != 0
    counts = nz.sum(dim=1, dtype=torch.int32)
    indptr = pipeline.exclusive_scan(counts.contiguous())
    cols = nz.nonzero()[:, 1].to(torch.int32).contiguous()
    return pipeline.DeviceCSR(indptr, cols, K[nz].contiguous(), tuple(K.shape))


def _one_hot(labels, L):
    C = torch.zeros((labels.shape[0], L), dtype=torch.float64, device=labels.device)
    C[torch.arange(labels.shape[0], device=labels.device), labels.long()] = 1.0
    return C


def dense_landmark(K, labels, L):
    """Landmark operator of a DENSE kernel (graphs.py:1169-1246 on an ndarray K): two plain float64 library GEMMs --
    pmn = C^T K, op = rownorm(pmn) rownorm(pmn^T).  (Going through the sparse aggregation kernels would turn every
    row of the n x n matrix into a "hub row".)  Plumbing, not a hot path: dense graphs are the small-n case."""
    C = _one_hot(labels, L)
    pmn = C.T @ K                                   # [L, n]: sum of the kernel rows of each cluster
    pnm = pmn.T.contiguous()                        # [n, L]
    rs_m = pmn.abs().sum(1, keepdim=True)
    rs_n = pnm.abs().sum(1, keepdim=True)
    pmn_n = torch.where(rs_m != 0, pmn / rs_m, pmn)
    pnm_n = torch.where(rs_n != 0, pnm / rs_n, pnm)
    op = pmn_n @ pnm_n
    return op.cpu().numpy(), pnm_n.cpu().numpy()


def dense_landmark_extend(Kyx, labels, L):
    """rownorm(Kyx C) for a dense out-of-sample kernel (graphs.py:1272-1288)."""
    agg = Kyx @ _one_hot(labels, L)
    rs = agg.abs().sum(1, keepdim=True)
    return torch.where(rs != 0, agg / rs, agg).cpu().numpy()
