"""Measurement of the SURVEY section 8f rows built so far, to the same bar as the hot path: device time (CUDA events),
algorithmic bytes against the measured HBM peak for the HBM-bound kernel, the host (reference) arithmetic timed
beside it on a bounded sample, and a parity figure.  One JSON line per row on stdout.

  f2/f3  CSR x dense product (csrc/spmm.cu): diffusion step P . X, interpolation T . E
  f1     spectral landmark selection on the device (randomized SVD + mini-batch k-means)
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphtools_b200 as gt
from graphtools_b200 import pipeline, spectral, synth

warnings.simplefilter("ignore")


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6551.7), "measured"
    return 6650.0, "fallback"


def ev_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--f", type=int, default=100)
    ap.add_argument("--n-host", type=int, default=100_000, help="size of the host-side (scipy / sklearn) comparison")
    a = ap.parse_args()
    hbm, how = peaks()
    X, _ = synth.gaussian_mixture(a.n, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    G = gt.Graph(torch.from_numpy(X).cuda(), knn=5, decay=40, verbose=0)
    K, Pv = G._dev_kernel, G._dev_P
    n, nnz = K.shape[0], K.nnz

    # ---- f3: one diffusion step P . S  (S = n x f float64)
    S = torch.randn((n, a.f), dtype=torch.float64, device="cuda")
    ms, out = ev_time(lambda: pipeline.spmm(K, S, Pv))
    bytes_alg = nnz * (12 + 8 * a.f) + 8 * n * a.f + 8 * (n + 1)
    P_host = K.to_scipy(Pv)
    S_host = S.cpu().numpy()
    m = min(n, a.n_host)
    t0 = time.perf_counter()
    ref = P_host[:m].dot(S_host)
    t_host = time.perf_counter() - t0
    same = bool(np.array_equal(out[:m].cpu().numpy(), ref))
    print(json.dumps({"row": "f3 diffusion step diff_op . S (gtb_spmm_csr)", "n": n, "nnz": nnz, "f": a.f,
                      "device_ms": ms, "algorithmic_GB": bytes_alg / 1e9, "achieved_GBps": bytes_alg / ms / 1e6,
                      "hbm_peak_GBps": hbm, "peak_source": how, "frac": bytes_alg / ms / 1e6 / hbm,
                      "host_scipy_s_for_rows": t_host, "host_rows": m,
                      "host_scipy_extrapolated_s": t_host * n / m, "bit_identical_to_scipy": same}), flush=True)

    # ---- f2: out-of-sample interpolation of an embedding: extend_to_data(Y) . E, device resident
    ny = 100_000
    Y = X[:: n // ny][:ny] + np.float32(0.01)
    Eemb = np.random.default_rng(0).standard_normal((n, 10))
    Yd = torch.from_numpy(Y).cuda()
    t0 = time.perf_counter()
    got = G.interpolate(Eemb, Y=Yd)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    T = G.extend_to_data(Yd)
    t0 = time.perf_counter()
    ref = T.dot(Eemb)
    t_dot = time.perf_counter() - t0
    print(json.dumps({"row": "f2 interpolate(E, Y): out-of-sample kernel + normalise + product, device resident",
                      "n": n, "n_y": ny, "f": 10, "device_s_incl_h2d_of_E_and_d2h": t_dev,
                      "host_scipy_dot_only_s": t_dot, "bit_identical_to_scipy": bool(np.array_equal(got, ref))}),
          flush=True)
    del S, out, got, T
    torch.cuda.empty_cache()

    # ---- f1: spectral landmarks on the device at full size; the host path (reference) at n_host
    L = 2000
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    s, Vt = spectral.randomized_svd_vt(K, G._dev_degree, 100, random_state=42)
    e1.record()
    feats = pipeline.spmm(K, Vt.T.contiguous(), Pv)
    labels, centers, inertia = spectral.minibatch_kmeans(feats, L, init_size=3 * L, batch_size=10000, random_state=42)
    e2.record()
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    svd_ms, km_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    # host path on a bounded size
    from sklearn.cluster import MiniBatchKMeans
    from sklearn.utils.extmath import randomized_svd
    Xs, _ = synth.gaussian_mixture(a.n_host, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    Gs = gt.Graph(Xs, knn=5, decay=40, verbose=0)
    A, Ph = Gs.diff_aff, Gs.diff_op
    Ls = max(50, L * a.n_host // n)
    t0 = time.perf_counter()
    _, _, VT = randomized_svd(A, n_components=100, random_state=42)
    t_svd = time.perf_counter() - t0
    F = Ph.dot(VT.T)
    t0 = time.perf_counter()
    km = MiniBatchKMeans(Ls, init_size=3 * Ls, n_init=1, batch_size=10000, random_state=42).fit(F)
    t_km = time.perf_counter() - t0
    ls_dev, _, inertia_dev = spectral.minibatch_kmeans(torch.from_numpy(F).cuda(), Ls, init_size=3 * Ls, batch_size=10000,
                                                       random_state=42)
    print(json.dumps({"row": "f1 spectral landmark selection (randomized SVD n_svd=100 + MiniBatchKMeans)",
                      "n": n, "n_landmark": L, "device_svd_ms": svd_ms, "device_kmeans_ms": km_ms,
                      "device_total_s": t_dev, "n_clusters_used": int(torch.unique(labels).numel()),
                      "host_n": a.n_host, "host_n_landmark": Ls, "host_sklearn_svd_s": t_svd,
                      "host_sklearn_kmeans_s": t_km,
                      "kmeans_inertia_device_over_sklearn_same_features": inertia_dev / km.inertia_}), flush=True)


if __name__ == "__main__":
    main()
