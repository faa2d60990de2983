"""K9 (csrc/gemm.cu): float64-faithful dense products on the int8 tensor cores and the landmark diffusion chain
landmark_op^t (the callers' np.linalg.matrix_power on the operator of graphs.py:1240-1243).

Oracle = numpy: exact integer products where the operands are integers (bit-exact: the digit planes ARE the numbers),
extended-precision products for general float64 operands, np.linalg.matrix_power for the chain."""
import numpy as np
import pytest
import torch

import graphtools_b200 as gt
from graphtools_b200 import dense, pipeline, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("row_bytes", [32, 64, 128])
@pytest.mark.parametrize("shape", [(128, 64, 128), (200, 150, 300), (257, 513, 129), (1000, 70, 2000)])
def test_gemm_integer_operands_exact(row_bytes, shape):
    """Integer-valued operands are represented exactly by the digit planes and accumulated exactly in int32, so the
    product must equal numpy's bit for bit -- any mistake in a descriptor, swizzle or digit-pair schedule shows here."""
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.integers(-30000, 30000, size=(M, K)).astype(np.float64)
    B = rng.integers(-30000, 30000, size=(K, N)).astype(np.float64)
    A[rng.random(A.shape) < 0.3] = 0
    C = dense.gemm_f64(pipeline.to_device(A), pipeline.to_device(B), row_bytes=row_bytes).cpu().numpy()
    assert np.array_equal(C, A @ B)


@pytest.mark.parametrize("slices", [4, 7])
def test_gemm_f64_error_bound(slices):
    """General float64 operands with a wide dynamic range, against an extended-precision product: the error stays
    under the fixed-point truncation bound (S + 2) K 2^(-8S - 2) sa_i sb_j (+ one float64 rounding per order)."""
    rng = np.random.default_rng(5)
    M, N, K = 300, 260, 1100
    A = rng.normal(size=(M, K)) * np.exp(4 * rng.normal(size=(M, K)))
    B = rng.normal(size=(K, N)) * np.exp(4 * rng.normal(size=(K, N)))
    C = dense.gemm_f64(pipeline.to_device(A), pipeline.to_device(B), slices=slices).cpu().numpy()
    R = (A.astype(np.longdouble) @ B.astype(np.longdouble))
    sa = 4 * 2.0 ** np.ceil(np.log2(np.abs(A).max(1)))[:, None]      # >= the kernel's row / column scales
    sb = 4 * 2.0 ** np.ceil(np.log2(np.abs(B).max(0)))[None, :]
    bound = (slices + 2) * K * 2.0 ** (-8 * slices - 2) * sa * sb + 8 * 2.0 ** -53 * np.abs(R).astype(np.float64)
    err = np.abs(C.astype(np.longdouble) - R).astype(np.float64)
    assert (err <= bound).all(), float((err / bound).max())


def test_gemm_row_stochastic_and_zero_rows():
    rng = np.random.default_rng(1)
    P = rng.random((500, 500)) ** 6
    P[7] = 0
    P[:, 11] = 0
    rs = P.sum(1, keepdims=True)
    P = np.where(rs > 0, P / np.where(rs > 0, rs, 1), 0)
    Q = dense.gemm_f64(pipeline.to_device(P), pipeline.to_device(P)).cpu().numpy()
    R = P @ P
    assert np.allclose(Q, R, rtol=1e-12, atol=1e-18)
    assert (Q[7] == 0).all() and (Q[:, 11] == 0).all()


@pytest.mark.parametrize("t", [0, 1, 2, 3, 5, 16, 37, 100])
def test_matrix_power_matches_numpy(t):
    rng = np.random.default_rng(2)
    P = rng.random((700, 700)) ** 10
    P /= P.sum(1, keepdims=True)
    got = dense.matrix_power(pipeline.to_device(P), t).cpu().numpy()
    ref = np.linalg.matrix_power(P, t)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-15), float(np.abs(got - ref).max())


def test_landmark_op_power_c5_shape():
    """The chain on a real landmark operator (100k x 100, L = 2000, random landmarking): landmark_op^t against
    np.linalg.matrix_power of the operator the graph returns, rtol 1e-5; rows stay stochastic."""
    X, _ = synth.gaussian_mixture(100_000, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, n_landmark=2000, random_landmarking=True, random_state=42, verbose=0)
    op = G.landmark_op
    assert op.shape == (2000, 2000)
    for t in (2, 8, 33):
        got = G.landmark_op_power(t)
        ref = np.linalg.matrix_power(op, t)
        # entries below 1e-8 of the unit row mass are compared absolutely: the digit planes are fixed point relative
        # to the row / column maximum (error <= 1e-15 of it), numpy's non-negative dot products componentwise
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-13), (t, float(np.abs(got - ref).max()))
        assert np.allclose(got.sum(1), 1.0, rtol=0, atol=1e-12)
    dev = G.landmark_op_power(8, return_device=True)
    assert isinstance(dev, torch.Tensor) and dev.dtype == torch.float64


def test_dense_graph_diffuse_uses_exact_products():
    """diffuse() on a dense (exact) graph: t products with the device-resident P, against the host products."""
    X, _ = synth.gaussian_mixture(600, 20, n_clusters=4, intrinsic_dim=5, seed=9)
    G = gt.Graph(X, graphtype="exact", knn=5, decay=20, thresh=0, verbose=0)
    sig = np.random.default_rng(0).normal(size=(600, 3))
    got = G.diffuse(sig, t=3)
    P = G.diff_op
    ref = P @ (P @ (P @ sig))
    assert np.allclose(got, ref, rtol=1e-11, atol=1e-14)
