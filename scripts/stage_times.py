"""Per-entry-point CUDA-event timings of one kNN graph build (profiling aid, not the bench)."""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import graphtools_b200 as gt
from graphtools_b200 import _engine as E, pipeline, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--q", type=int, default=10)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
X, _ = synth.gaussian_mixture(a.n, a.d, n_clusters=20, intrinsic_dim=(a.q if a.q > 0 else None), seed=0)
for rep in range(a.reps):
    E.timing = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    G = gt.Graph(X, knn=5, decay=40, thresh=1e-4, verbose=0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    K = G.kernel; P = G.diff_op
    t2 = time.perf_counter()
    tm = E.timings_ms()
    E.timing = None
    print("rep %d n=%d d=%d q=%d: device build %.1f ms, +materialise %.1f ms, stats %s" % (
        rep, a.n, a.d, a.q, (t1 - t0) * 1e3, (t2 - t1) * 1e3, pipeline.stats()))
    for k, (c, ms) in sorted(tm.items(), key=lambda kv: -kv[1][1]):
        print("   %-24s calls=%d %.3f ms" % (k, c, ms))
    flop = 2.0 * a.n * a.n * a.d
    key = next(k for k in ("gtb_knn_topk_tc_seeded", "gtb_knn_topk_tc", "gtb_knn_topk_simt") if k in tm)
    ms = tm[key][1]
    print("   %s: %.2f TFLOP/s algorithmic (2*N*N*d)" % (key, flop / ms / 1e9))
