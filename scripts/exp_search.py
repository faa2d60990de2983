"""Times the tensor-core search kernel alone (gtb_knn_topk_tc) on the headline shape -- kernel experiments.
    GTB_LIB=path/to/variant.so python scripts/exp_search.py [--n 1000000] [--d 100] [--dtype 1] [--reps 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphtools_b200 import _engine as E, pipeline, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--dtype", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tag", default="")
a = ap.parse_args()
X, _ = synth.gaussian_mixture(a.n, a.d, n_clusters=50, intrinsic_dim=10, seed=3)
ref = pipeline.SearchOperand(torch.from_numpy(X).cuda())
scale = pipeline.fp16_scale(ref.norm_max()) if a.dtype >= 2 else 1.0
q_hi, q_lo, q_n2 = ref.tc(0, a.dtype, scale)
r_hi, r_lo, _ = ref.tc(1, a.dtype, scale)
q_n2 = q_n2 * (scale * scale)
Kp = ref.kp(a.dtype)
pace = torch.zeros(1, dtype=torch.int32, device="cuda") if int(os.environ.get("GTB_TC_PACING", "1")) else None
cluster = int(os.environ.get("GTB_TC_CLUSTER", "2"))
ls = int(os.environ.get("GTB_TC_LIST", "32"))
qtiles = int(os.environ.get("GTB_TC_QTILES", "1"))
cand = torch.empty((a.n, 2 * ls), dtype=torch.int32, device="cuda")
tau = torch.empty((a.n, 2), dtype=torch.float32, device="cuda")
scratch = torch.empty((E.lib().gtb_tc_scratch_bytes(ref.n_pad),), dtype=torch.uint8, device="cuda")
times = []
stride = int(os.environ.get("GTB_TC_SEED_STRIDE", "16"))
seed_ms = []
for rep in range(a.reps + 1):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    seed = None
    if a.dtype == 3 and stride > 1:
        seed = torch.empty((a.n, 2), dtype=torch.float32, device="cuda")
        E.call("gtb_knn_seed_tc", q_hi, q_n2, a.n, ref.n_pad, r_hi, a.n, ref.n_pad, Kp, cluster, stride, seed, pace)
    e1.record()
    E.call("gtb_knn_topk_tc_seeded", q_hi, q_lo, q_n2, a.n, ref.n_pad, r_hi, r_lo, a.n, ref.n_pad, Kp, a.dtype, ls, cluster,
           qtiles, seed, 1, cand, scratch, tau, pace)
    e2.record()
    torch.cuda.synchronize()
    times.append(e1.elapsed_time(e2))
    seed_ms.append(e0.elapsed_time(e1))
best = min(times[1:])
if a.dtype == 3:
    full = (cand >= 0).sum(1)
    print("EXP seed stride %d: seed pass %.1f ms; lists full %.3f, mean candidates %.1f, tau finite %.4f" % (
        stride, min(seed_ms[1:]), float((full == cand.shape[1]).float().mean()), float(full.float().mean()),
        float(torch.isfinite(tau[:, 0]).float().mean())), flush=True)
print("EXP %s qtiles=%d list=%d lib=%s n=%d d=%d Kp=%d dtype=%d: %s ms (best %.1f) -> %.1f TFLOP/s algorithmic; cand checksum %d" % (
    a.tag, qtiles, ls, os.path.basename(E.LIB_PATH), a.n, a.d, Kp, a.dtype, ["%.1f" % t for t in times], best,
    2.0 * a.n * a.n * a.d / best / 1e9, int(cand.clamp(min=0).to(torch.int64).sum().item())), flush=True)
