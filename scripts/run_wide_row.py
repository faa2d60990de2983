"""Wide data (d + 2 > 128) on the tensor cores vs the CUDA-core search it used to fall back to: device time of the
kNN graph build (kernel + diff_op) at n x d, once per search implementation, with the result compared bit for bit."""
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
from graphtools_b200 import pipeline, synth

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200_000)
ap.add_argument("--d", type=int, default=300)
a = ap.parse_args()
X, _ = synth.gaussian_mixture(a.n, a.d, n_clusters=20, intrinsic_dim=10, seed=5)
Xd = torch.from_numpy(X).cuda()
out = {}
for impl in ("auto", "simt"):
    os.environ["GTB_SEARCH_IMPL"] = impl
    best = 1e30
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        G = gt.Graph(Xd, knn=5, decay=40, thresh=1e-4, verbose=0)
        G._ensure_built()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    st = pipeline.stats()
    out[impl] = {"impl_used": st["impl"], "device_s": best, "points_per_s": a.n / best,
                 "algorithmic_TFLOPs": 2.0 * a.n * a.n * a.d / best / 1e12, "radius_rows": st["radius_rows"],
                 "K": G.kernel}
same = (out["auto"]["K"] != out["simt"]["K"]).nnz == 0
for v in out.values():
    v.pop("K")
print(json.dumps({"row": "wide data on tensor cores (chunked fp16x2 sweep)", "n": a.n, "d": a.d, **out,
                  "bit_identical": bool(same), "speedup": out["simt"]["device_s"] / out["auto"]["device_s"]}))
