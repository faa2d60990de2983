"""Synthetic inputs for parity tests and bench.py.

All generators return **float32** arrays: the CUDA path consumes float32 data and
the CPU oracle is handed ``X.astype(np.float64)`` so both sides see bit-identical
inputs (input rounding alone would otherwise break rtol 1e-5 at
decay=40).
"""
import numpy as np


def gaussian_mixture(n, d, n_clusters=20, intrinsic_dim=10, seed=0, noise=0.05,
                     center_scale=5.0, dtype=np.float32):
    """Mixture of low-intrinsic-dimension Gaussians embedded in ``d`` dims.

    ``X = centers[label] + z @ B[label] + noise * eps`` with ``z ~ N(0, I_q)``,
    ``B ~ N(0,1)/sqrt(q)``; ``intrinsic_dim=None`` (or >= d) gives the isotropic
    stress variant ``X = centers[label] + eps``.
    """
    rng = np.random.default_rng(seed)
    centers = rng.normal(0.0, center_scale, size=(n_clusters, d))
    labels = rng.integers(0, n_clusters, size=n)
    if intrinsic_dim is None or intrinsic_dim >= d:
        X = centers[labels] + rng.normal(size=(n, d))
        return np.ascontiguousarray(X.astype(dtype)), labels
    q = int(intrinsic_dim)
    basis = rng.normal(size=(n_clusters, q, d)) / np.sqrt(q)
    X = np.empty((n, d), dtype=np.float64)
    # chunked so the 1M x 100 headline shape never needs an [n, q, d] temporary
    step = 65536
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        lab = labels[lo:hi]
        z = rng.normal(size=(hi - lo, q))
        eps = rng.normal(size=(hi - lo, d))
        part = centers[lab] + noise * eps
        order = np.argsort(lab, kind="stable")
        lab_sorted = lab[order]
        starts = np.searchsorted(lab_sorted, np.arange(n_clusters), side="left")
        ends = np.searchsorted(lab_sorted, np.arange(n_clusters), side="right")
        for c in range(n_clusters):
            rows = order[starts[c]:ends[c]]
            if rows.size:
                part[rows] += z[rows] @ basis[c]
        X[lo:hi] = part
    return np.ascontiguousarray(X.astype(dtype)), labels


def batched_mixture(n_per_batch, n_batches, d, n_clusters=20, intrinsic_dim=10, seed=2,
                    batch_shift=0.5, dtype=np.float32):
    """Config-4 style input: ``n_batches`` contiguous batches, each the same mixture
    shifted by a per-batch N(0, batch_shift^2) offset.  Returns (X, sample_idx)."""
    n = n_per_batch * n_batches
    X, _ = gaussian_mixture(n, d, n_clusters, intrinsic_dim, seed, dtype=np.float64)
    rng = np.random.default_rng(seed + 1000003)
    shifts = rng.normal(0.0, batch_shift, size=(n_batches, d))
    sample_idx = np.repeat(np.arange(n_batches), n_per_batch)
    X = X + shifts[sample_idx]
    return np.ascontiguousarray(X.astype(dtype)), sample_idx


def swiss_roll(n=1000, noise=0.5, seed=42, dtype=np.float32):
    """Two-batch swiss roll in the spirit of the reference's test fixture
    (reference test/load_tests/__init__.py:102-114); own formulation."""
    rng = np.random.default_rng(seed)
    t = 1.5 * np.pi * (1 + 2 * rng.random(n))
    height = 21 * rng.random(n)
    X = np.stack([t * np.cos(t), height, t * np.sin(t)], axis=1)
    X += noise * rng.normal(size=X.shape)
    sample_idx = rng.integers(0, 2, size=n)
    return np.ascontiguousarray(X.astype(dtype)), sample_idx
