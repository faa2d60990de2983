"""Parity comparator.

Structure must match exactly except for entries that sit on a discontinuity of the reference's own
definition: |w - thresh| / thresh < 1e-4 (threshold cut) -- those are counted and must be a
vanishing fraction of nnz.  Values on the common structure: |gpu - ref| <= rtol * |ref|, rtol 1e-5
(north_star tolerance for kernel and diff_op values).
"""
import numpy as np
from scipy import sparse

RTOL = 1e-5


def _coo_keys(M):
    M = sparse.coo_matrix(M)
    return M.row.astype(np.int64) * M.shape[1] + M.col.astype(np.int64), M.data


def compare_sparse(gpu, ref, rtol=RTOL, thresh=None, max_exempt_frac=1e-5, atol=0.0, what="matrix"):
    """Returns dict(n_exempt, max_rel). Raises AssertionError on a parity violation."""
    assert gpu.shape == ref.shape, (gpu.shape, ref.shape)
    gpu = sparse.csr_matrix(gpu); ref = sparse.csr_matrix(ref)
    gpu.sum_duplicates(); ref.sum_duplicates()
    kg, vg = _coo_keys(gpu)
    kr, vr = _coo_keys(ref)
    og, orr = np.argsort(kg), np.argsort(kr)
    kg, vg, kr, vr = kg[og], vg[og], kr[orr], vr[orr]
    common, ig, ir = np.intersect1d(kg, kr, assume_unique=True, return_indices=True)
    only_g = np.setdiff1d(np.arange(len(kg)), ig, assume_unique=True)
    only_r = np.setdiff1d(np.arange(len(kr)), ir, assume_unique=True)
    n_exempt = 0
    if len(only_g) or len(only_r):
        assert thresh is not None and thresh > 0, "{}: structure differs ({} extra, {} missing)".format(
            what, len(only_g), len(only_r))
        stray = np.concatenate([vg[only_g], vr[only_r]])
        bad = np.abs(stray - thresh) / thresh >= 1e-4
        assert not bad.any(), "{}: {} structural differences away from the threshold (e.g. value {})".format(
            what, int(bad.sum()), stray[bad][0])
        n_exempt = len(stray)
        assert n_exempt <= max(1, max_exempt_frac * ref.nnz), "{}: too many threshold-boundary entries ({})".format(
            what, n_exempt)
    a, b = vg[ig], vr[ir]
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol
    worst = float((err / np.maximum(np.abs(b), 1e-300)).max()) if len(b) else 0.0
    assert (err <= tol).all(), "{}: max relative error {:.3e} exceeds rtol {:g}".format(what, worst, rtol)
    return {"n_exempt": n_exempt, "max_rel": worst}


def compare_dense(gpu, ref, rtol=RTOL, thresh=None, what="matrix"):
    gpu = np.asarray(gpu); ref = np.asarray(ref)
    assert gpu.shape == ref.shape
    err = np.abs(gpu - ref)
    ok = err <= rtol * np.abs(ref)
    if thresh:
        # entries zeroed by the threshold on one side only
        edge = (np.abs(np.maximum(gpu, ref) - thresh) / thresh < 1e-4) & ((gpu == 0) | (ref == 0))
        ok |= edge
    worst = float((err[ok] / np.maximum(np.abs(ref[ok]), 1e-300)).max()) if ok.any() else 0.0
    assert ok.all(), "{}: {} entries outside rtol {:g} (max abs err {:.3e})".format(what, int((~ok).sum()), rtol,
                                                                                    float(err[~ok].max()))
    return {"max_rel": worst}
