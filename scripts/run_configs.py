"""Runs the five BASELINE.json configurations on one B200 and records device timings, a parity check
against the CPU oracle at a size the oracle finishes in seconds, and the oracle's own wall time
(host cores) beside them.  Writes one JSON line per config to stdout (profiles/rNN_configs.jsonl)."""
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
from graphtools_b200 import _engine as E, pipeline, synth
from oracle import graph_oracle as go
from tests.parity import compare_dense, compare_sparse

warnings.simplefilter("ignore")


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def cpu_time(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    from sklearn.datasets import load_digits
    X = load_digits().data.astype(np.float32)
    t_gpu, G = timed(lambda: _kd(gt.Graph(X, knn=5, decay=40, verbose=0)), reps=3)
    t_cpu, (K, P) = cpu_time(lambda: go.knn_graph(X.astype(np.float64), knn=5, decay=40))
    r = compare_sparse(G.kernel, K, thresh=1e-4)
    compare_sparse(G.diff_op, P)
    emit(config="C1 digits 1797x64 kNN knn=5 decay=40", gpu_s=t_gpu, cpu_oracle_s=t_cpu, nnz=int(K.nnz),
         max_rel_err=r["max_rel"], exempt=r["n_exempt"], e2e=True)


def _kd(G):
    G.kernel, G.diff_op
    return G


def _device_build(X, **kw):
    G = gt.Graph(X, verbose=0, **kw)
    G._ensure_built()
    return G


def c2(n=100_000):
    X, _ = synth.gaussian_mixture(n, 100, n_clusters=20, intrinsic_dim=10, seed=0)
    Xd = torch.from_numpy(X).cuda()
    t_dev, G = timed(lambda: _device_build(Xd, knn=5, decay=40, thresh=1e-4))
    t_e2e, G2 = timed(lambda: _kd(gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)))
    t_cpu, (K, P) = cpu_time(lambda: go.knn_graph(X.astype(np.float64), knn=5, decay=40, thresh=1e-4))
    r = compare_sparse(G2.kernel, K, thresh=1e-4)
    compare_sparse(G2.diff_op, P, thresh=1e-4 if r["n_exempt"] else None, rtol=1e-5 if not r["n_exempt"] else 1e-3)
    emit(config="C2 mixture %dx100 kNN knn=5 decay=40 thresh=1e-4" % n, gpu_device_s=t_dev, gpu_e2e_s=t_e2e,
         cpu_oracle_s=t_cpu, points_per_s_device=n / t_dev, points_per_s_cpu=n / t_cpu, nnz=int(K.nnz),
         max_rel_err=r["max_rel"], exempt=r["n_exempt"], stats=pipeline.stats())


def c3(n=50_000, n_par=4000):
    X, _ = synth.gaussian_mixture(n, 50, n_clusters=10, intrinsic_dim=10, seed=1)
    Xs = X[:n_par]
    t_cpu, K = cpu_time(lambda: go.finish_kernel(go.exact_kernel(Xs.astype(np.float64), knn=5, decay=40, thresh=1e-4)))
    P = go.diff_op(K)
    Gs = _kd(gt.Graph(Xs, graphtype="exact", knn=5, decay=40, thresh=1e-4, verbose=0))
    r = compare_dense(Gs.kernel, K, thresh=1e-4)
    compare_dense(Gs.diff_op, P, thresh=1e-4)
    del Gs
    torch.cuda.empty_cache()
    Xd = torch.from_numpy(X).cuda()
    E.timing = {}
    t_dev, G = timed(lambda: _device_build(Xd, graphtype="exact", knn=5, decay=40, thresh=1e-4), reps=1)
    tm = E.timings_ms(); E.timing = None
    stage = {k: v[1] / 2 for k, v in tm.items()}          # warm-up + timed run were both recorded
    dens_ms = stage.get("gtb_csr_to_dense")
    nnz_frac = float((G._dev_kernel != 0).sum().item()) / float(n * n)
    emit(config="C3 exact %dx50 knn=5 decay=40 (sparse tensor-core route + densify)" % n, gpu_device_s=t_dev,
         stage_ms=stage, densify_ms_K_and_P=dens_ms,
         hbm_write_GBps_densify=(16.0 * n * n / dens_ms / 1e6) if dens_ms else None, nonzero_fraction=nnz_frac,
         parity_n=n_par, cpu_oracle_s_at_parity_n=t_cpu, cpu_extrapolated_s=t_cpu * (n / n_par) ** 2,
         max_rel_err=r["max_rel"])


def c4(n_per=250_000, n_par=3000):
    Xs, idx_s = synth.batched_mixture(n_par, 4, 100, n_clusters=10, intrinsic_dim=10, seed=2)
    t_cpu, R = cpu_time(lambda: go.mnn_kernel(Xs.astype(np.float64), idx_s, knn=5, decay=40, thresh=1e-4))
    K = go.finish_kernel(R, "mnn", 0.5)
    Gs = _kd(gt.Graph(Xs, sample_idx=idx_s, kernel_symm="mnn", theta=0.5, knn=5, decay=40, verbose=0))
    r = compare_sparse(Gs.kernel, K, thresh=1e-4)
    del Gs
    X, idx = synth.batched_mixture(n_per, 4, 100, n_clusters=20, intrinsic_dim=10, seed=2)
    Xd = torch.from_numpy(X).cuda()
    t_dev, G = timed(lambda: _device_build(Xd, sample_idx=idx, kernel_symm="mnn", theta=0.5, knn=5, decay=40), reps=1)
    n = 4 * n_per
    emit(config="C4 MNN 4x%dx100 theta=0.5 knn=5 decay=40" % n_per, gpu_device_s=t_dev, points_per_s_device=n / t_dev,
         nnz=int(G._dev_kernel.nnz), parity_n=4 * n_par, cpu_oracle_s_at_parity_n=t_cpu,
         cpu_extrapolated_s=t_cpu * (n_per / n_par) ** 2, max_rel_err=r["max_rel"], exempt=r["n_exempt"])


def c5(n=1_000_000, n_par=20_000, L=2000):
    Xs, _ = synth.gaussian_mixture(n_par, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    K, P = go.knn_graph(Xs.astype(np.float64), knn=5, decay=40)
    t_cl, clusters = cpu_time(lambda: go.random_landmark_clusters(Xs.astype(np.float64), 500, 42))
    t_cpu, (op, pnm) = cpu_time(lambda: go.landmark_operator(K, clusters))
    Gs = gt.Graph(Xs, knn=5, decay=40, n_landmark=500, random_landmarking=True, random_state=42, verbose=0)
    agree = float(np.mean(Gs.clusters == clusters))
    Gs.clusters = clusters
    r = compare_dense(Gs.landmark_op, op)
    compare_sparse(Gs.transitions, pnm)
    del Gs
    X, _ = synth.gaussian_mixture(n, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    Xd = torch.from_numpy(X).cuda()

    def build():
        G = gt.Graph(Xd, knn=5, decay=40, n_landmark=L, random_landmarking=True, random_state=42, verbose=0)
        G.build_landmark_op()
        return G
    E.timing = {}
    t_dev, G = timed(build, reps=1)
    tm = E.timings_ms(); E.timing = None
    lm = {k: v[1] / 2 for k, v in tm.items() if "cluster" in k or "landmark" in k}
    emit(config="C5 landmark %dx100 knn=5 decay=40 n_landmark=%d (random landmarking on GPU)" % (n, L),
         gpu_device_s_kernel_plus_landmark_op=t_dev, landmark_kernel_ms=lm, L_eff=int(G.landmark_op.shape[0]),
         parity_n=n_par, cluster_agreement_at_parity_n=agree, cpu_oracle_landmark_s_at_parity_n=t_cpu,
         cpu_oracle_cluster_s_at_parity_n=t_cl, max_rel_err_landmark_op=r["max_rel"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c2,c3,c4,c5")
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke)")
    a = ap.parse_args()
    todo = a.only.split(",")
    for name in todo:
        try:
            if a.small:
                {"c1": c1, "c2": lambda: c2(20000), "c3": lambda: c3(6000, 2000), "c4": lambda: c4(20000, 1500),
                 "c5": lambda: c5(100000, 10000, 500)}[name]()
            else:
                {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[name]()
        except Exception as ex:  # keep going: one JSON line per config either way
            import traceback
            emit(config=name, error=repr(ex), trace=traceback.format_exc()[-1500:])
        torch.cuda.empty_cache()
