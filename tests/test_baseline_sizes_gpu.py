"""Parity at the BASELINE.json sizes: every config is compared with the CPU oracle -- in full
where the oracle finishes in seconds (C2), on sampled rows against the FULL reference set where it does not
(headline 1M, C3 50k exact), at a reduced batch size for the MNN config (C4) -- plus the certification bound of
the tensor-core search at 1M.  All CUDA work goes through the public Graph API / the C ABI.

reference: graphs.py:819-982 (kNN kernel rows), :1546-1609 (exact), :1857-1936 (MNN), base.py:557-646."""
import warnings

import numpy as np
import pytest
from scipy import sparse

import graphtools_b200 as gt
from graphtools_b200 import pipeline, synth
from tests.parity import compare_dense, compare_sparse

pytestmark = pytest.mark.gpu

KNN, DECAY, THRESH = 5, 40, 1e-4


@pytest.fixture(scope="module")
def headline():
    """The headline workload (bench.py): 1M x 100 mixture, 50 clusters, seed 3 -- generated once per module."""
    X, _ = synth.gaussian_mixture(1_000_000, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    return X


def test_c2_full_oracle_100k():
    """(a) BASELINE config 2 in full: K and P of the 100k x 100 kNN graph against the oracle."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(100_000, 100, n_clusters=20, intrinsic_dim=10, seed=0)
    K_ref, P_ref = go.knn_graph(X.astype(np.float64), knn=KNN, decay=DECAY, thresh=THRESH)
    G = gt.Graph(X, knn=KNN, decay=DECAY, thresh=THRESH, verbose=0)
    r = compare_sparse(G.kernel, K_ref, thresh=THRESH, what="C2.K")
    assert r["max_rel"] < 1e-9
    if r["n_exempt"] == 0:
        assert np.array_equal(G.kernel.indices, K_ref.indices) and np.array_equal(G.kernel.indptr, K_ref.indptr)
        compare_sparse(G.diff_op, P_ref, what="C2.P")
    else:
        compare_sparse(G.diff_op, P_ref, thresh=THRESH, rtol=1e-3, what="C2.P")
    assert np.allclose(G.kernel_degree, go.kernel_degree(K_ref), rtol=1e-9)


def test_headline_sampled_rows_1m(headline):
    """(b) The headline build at full size: raw kernel rows of 4096 sampled queries against the oracle's rows for
    the same queries searched in the FULL 1M reference set (graphs.py:819-982), structure exact, values rtol 1e-5;
    then K = (R + R^T) / 2 and P on those rows from the GPU's own raw matrix (symmetrise + normalise parity)."""
    from oracle import graph_oracle as go
    X = headline
    n = X.shape[0]
    rows = np.unique(np.concatenate([np.linspace(0, n - 1, 4000).astype(np.int64), np.arange(96)]))
    X64 = X.astype(np.float64)
    g = go.KnnOracle(X64, knn=KNN, decay=DECAY, thresh=THRESH, n_jobs=-1)
    R_ref = g.kernel_to_data(X64[rows], knn=KNN + 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        G = gt.Graph(X, knn=KNN, decay=DECAY, thresh=THRESH, verbose=0, initialize=False)
        R = G.build_kernel().to_scipy()
        r = compare_sparse(R[rows], R_ref, thresh=THRESH, what="headline.R[rows]")
        assert r["n_exempt"] <= 2 and r["max_rel"] < 1e-9
        K, P = G.kernel, G.diff_op
    # the GPU's K against scipy's symmetrisation of the GPU's own raw kernel (bitwise), P against sklearn's rule
    K_ref = ((R + R.T) / 2).tocsr()
    K_ref.sort_indices()
    assert np.array_equal(K.indptr, K_ref.indptr) and np.array_equal(K.indices, K_ref.indices)
    assert np.array_equal(K.data, K_ref.data)
    P_ref = go.diff_op(K_ref)
    compare_sparse(P[rows], P_ref[rows], rtol=1e-12, what="headline.P[rows]")
    st = pipeline.stats()
    assert st["rows"] == n


def test_tc_certification_bound_1m(headline):
    """(c) The rounding bound the certification rests on (pipeline.eps_rel_tc16), at the headline size: for sampled
    queries, EVERY reference point that the tensor-core sweep did not return as a candidate has an exact squared
    distance >= tau - E."""
    import torch
    from graphtools_b200 import _engine as E
    X = headline
    n, d = X.shape
    Xd = pipeline.to_device_f32(X)
    ref = pipeline.SearchOperand(Xd)
    for tcd, eps_fn, qtiles in ((3, pipeline.eps_rel_tch1, 2), (1, pipeline.eps_rel_tc16, 1), (2, pipeline.eps_rel_tch, 2),
                                (2, pipeline.eps_rel_tch, 1), (0, pipeline.eps_rel_tc, 1)):
        ls = 32 if tcd == 3 else 16
        scale = pipeline.fp16_scale(ref.norm_max()) if tcd >= 2 else 1.0
        q_hi, q_lo, q_n2 = ref.tc(0, tcd, scale)
        r_hi, r_lo, _ = ref.tc(1, tcd, scale)
        cand = pipeline._empty((n, 2 * ls), torch.int32)
        tau = pipeline._empty((n, 2), torch.float32)
        scratch = pipeline._empty((E.lib().gtb_tc_scratch_bytes(ref.n_pad),), torch.uint8)
        pace = pipeline._empty((1,), torch.int32)
        seed = None
        if tcd == 3:
            # the one-product flavour as the pipeline runs it: thresholds seeded from every 16th reference tile
            seed = pipeline._empty((n, 2), torch.float32)
            E.call("gtb_knn_seed_tc", q_hi, q_n2 * (scale * scale), n, ref.n_pad, r_hi, n, ref.n_pad, ref.kp(tcd), 2, 16,
                   seed, pace)
        E.call("gtb_knn_topk_tc_seeded", q_hi, q_lo, q_n2 * (scale * scale), n, ref.n_pad, r_hi, r_lo, n, ref.n_pad,
               ref.kp(tcd), tcd, ls, 2, qtiles, seed, 1, cand, scratch, tau, pace)
        tau = tau / (scale * scale)
        del scratch
        rows = torch.linspace(0, n - 1, 192, device=Xd.device).long()
        X64d = Xd.double()
        Xc = X64d - ref.mean.double()                              # the centred rows the bound is stated on
        n2 = (Xc * Xc).sum(1)
        del Xc
        maxn2 = float(n2.max().item())
        assert maxn2 <= ref.maxnorm * (1 + 1e-6)
        worst = -np.inf
        for i in rows.tolist():
            d2 = ((X64d - X64d[i]) ** 2).sum(1)                    # exact float64 squared distances to all references
            is_cand = torch.zeros((n,), dtype=torch.bool, device=Xd.device)
            c = cand[i]
            is_cand[c[c >= 0].long()] = True
            t = float(tau[i].min().item())
            if not np.isfinite(t):
                continue
            Ebound = eps_fn(d) * (float(n2[i].item()) + maxn2)
            margin = float((d2[~is_cand] - (t - Ebound)).min().item())
            worst = max(worst, -margin / Ebound)
            assert margin >= 0.0, "row %d: a non-candidate lies %.3e inside tau - E (E = %.3e)" % (i, -margin, Ebound)
        print("dtype %d qtiles %d: worst used fraction of the bound %.3f" % (tcd, qtiles, worst))
        del cand, tau, X64d, n2


def test_c3_exact_sampled_rows_50k():
    """(d) BASELINE config 3 (exact graph, 50k x 50): sampled rows of K and P against a row-blocked restatement of
    graphs.py:1546-1609 + base.py:561,645.  Row i of K = (R[i, :] + R[:, i]) / 2 needs d(i, .) and every bandwidth;
    the bandwidths come from a blocked cdist + partition (the reference's own rule)."""
    import torch
    from scipy.spatial.distance import cdist
    n, d = 50_000, 50
    X, _ = synth.gaussian_mixture(n, d, n_clusters=10, intrinsic_dim=10, seed=1)
    X64 = X.astype(np.float64)
    bw = np.empty(n)
    for lo in range(0, n, 2500):
        D = cdist(X64[lo:lo + 2500], X64)
        bw[lo:lo + 2500] = np.max(np.partition(D, KNN + 1, axis=1)[:, :KNN + 1], axis=1)
    rows = np.linspace(0, n - 1, 64).astype(np.int64)
    D = cdist(X64[rows], X64)
    with np.errstate(invalid="ignore", divide="ignore"):
        R_rows = np.exp(-np.power(D / bw[rows][:, None], DECAY))      # R[i, :]
        R_cols = np.exp(-np.power(D / bw[None, :], DECAY))            # R[:, i]^T
    R_rows[R_rows < THRESH] = 0
    R_cols[R_cols < THRESH] = 0
    K_rows = (R_rows + R_cols) / 2
    G = gt.Graph(X, graphtype="exact", knn=KNN, decay=DECAY, thresh=THRESH, verbose=0)
    assert type(G).__name__ == "TraditionalGraph"
    G._ensure_built()
    Kd, Pd = G._dev_kernel, G._dev_P
    assert tuple(Kd.shape) == (n, n) and Kd.dtype == torch.float64
    idx = torch.from_numpy(rows).to(Kd.device)
    K_gpu = Kd[idx].cpu().numpy()
    assert np.array_equal(K_gpu != 0, K_rows != 0), "C3: support of the sampled rows differs"
    r = compare_dense(K_gpu, K_rows, what="C3.K[rows]")
    assert r["max_rel"] < 1e-9
    # P rows: K rows / row sums (the whole row is on hand, so the normaliser is the reference's)
    compare_dense(Pd[idx].cpu().numpy(), K_rows / K_rows.sum(1, keepdims=True), what="C3.P[rows]")
    assert torch.allclose(Pd.sum(1), torch.ones((n,), dtype=torch.float64, device=Pd.device), rtol=0, atol=1e-12)
    # symmetry of the full matrix, checked on the device (the 20 GB matrix never visits the host)
    blk = Kd[:4096]
    assert torch.equal(blk[:, :4096], blk[:, :4096].T)
    assert torch.equal(Kd[idx], Kd[:, idx].T.contiguous())


def test_c4_mnn_reduced_4x25k():
    """(e) BASELINE config 4 at 4 x 25k x 100: MNN kernel (16 searches, block assembly, mnn symmetrisation with
    theta = 0.5) against the COO restatement of graphs.py:1857-1936 (bit-identical to the reference at small n)."""
    from oracle import graph_oracle as go
    X, idx = synth.batched_mixture(25_000, 4, 100, n_clusters=20, intrinsic_dim=10, seed=2)
    R_ref = go.mnn_kernel(X.astype(np.float64), idx, knn=KNN, decay=DECAY, thresh=THRESH)
    K_ref = go.finish_kernel(R_ref, "mnn", 0.5)
    G = gt.Graph(X, sample_idx=idx, kernel_symm="mnn", theta=0.5, knn=KNN, decay=DECAY, thresh=THRESH, verbose=0)
    assert type(G).__name__ == "MNNGraph"
    r = compare_sparse(G.kernel, K_ref, thresh=THRESH, what="C4.K", max_exempt_frac=1e-5)
    if r["n_exempt"] == 0:
        compare_sparse(G.diff_op, go.diff_op(K_ref), what="C4.P")


def test_c5_landmark_operator_100k():
    """BASELINE config 5 at 100k x 100 with L = 2000: landmark_op / transitions against the oracle with the same
    clusters (random landmarking: clusters themselves bit-exact), and run-to-run bit reproducibility of the
    operator (deterministic segmented reductions, no floating-point atomics)."""
    from oracle import graph_oracle as go
    X, _ = synth.gaussian_mixture(100_000, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    K_ref, _ = go.knn_graph(X.astype(np.float64), knn=KNN, decay=DECAY, thresh=THRESH)
    clusters = go.random_landmark_clusters(X.astype(np.float64), 2000, 42)
    op_ref, pnm_ref = go.landmark_operator(K_ref, clusters)
    ops = []
    for _ in range(2):
        G = gt.Graph(X, knn=KNN, decay=DECAY, thresh=THRESH, n_landmark=2000, random_landmarking=True,
                     random_state=42, verbose=0)
        assert np.array_equal(G.clusters, clusters)
        ops.append(G.landmark_op.copy())
    assert np.array_equal(ops[0], ops[1]), "landmark_op differs between two runs"
    compare_dense(ops[0], op_ref, what="C5.landmark_op")
    compare_sparse(G.transitions, pnm_ref, what="C5.transitions")
