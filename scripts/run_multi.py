"""BASELINE configs C4 (MNN, rows sharded) and C5 (landmark, kNN sharded) on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/run_multi.py [--only c4,c5] [--small]

One process per GPU; every rank builds the complete graph (query rows sharded, reference set replicated, NCCL
all-gather / all-to-all as described in DESIGN.md section 6).  Times are CUDA-event milliseconds, max over ranks.
Rank 0 prints one JSON line per config with nnz and a checksum of K that must be identical for every N.
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphtools_b200 as gt
from graphtools_b200 import synth

warnings.simplefilter("ignore")


def timed(fn, reps=1):
    fn()
    best = 1e30
    for _ in range(reps):
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist.is_initialized():
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = min(best, float(ms.item()))
    return best, out


def checksum(K):
    return {"nnz": int(K.nnz), "sum": float(K.data.sum(dtype=torch.float64).item()),
            "isum": int(K.indices.to(torch.int64).sum().item())}


def c4(n_per):
    X, idx = synth.batched_mixture(n_per, 4, 100, n_clusters=20, intrinsic_dim=10, seed=2)
    Xd = torch.from_numpy(X).cuda()

    def build():
        G = gt.Graph(Xd, sample_idx=idx, kernel_symm="mnn", theta=0.5, knn=5, decay=40, verbose=0)
        G._ensure_built()
        return G
    ms, G = timed(build)
    n = 4 * n_per
    return dict(config="C4 MNN 4x%dx100 theta=0.5 knn=5 decay=40" % n_per, ms=ms, points_per_s=n / ms * 1e3,
                **checksum(G._dev_kernel))


def c5(n, L, spectral=False):
    X, _ = synth.gaussian_mixture(n, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    Xd = torch.from_numpy(X).cuda()
    if spectral:
        os.environ["GTB_SPECTRAL"] = "device"

    def build():
        G = gt.Graph(Xd, knn=5, decay=40, n_landmark=L, random_landmarking=not spectral, random_state=42, verbose=0)
        G.build_landmark_op()
        return G
    ms, G = timed(build)
    how = "spectral landmarks: device SVD + mini-batch k-means" if spectral else "random landmarking"
    return dict(config="C5 landmark %dx100 knn=5 decay=40 n_landmark=%d (%s)" % (n, L, how), ms=ms,
                L_eff=int(G.landmark_op.shape[0]),
                points_per_s=n / ms * 1e3, landmark_op_sum=float(np.sum(G.landmark_op)),
                landmark_op_trace=float(np.trace(G.landmark_op)), **checksum(G._dev_kernel))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c4,c5")
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    for name in a.only.split(","):
        if name == "c4":
            r = c4(20_000 if a.small else 250_000)
        else:
            r = c5(100_000 if a.small else 1_000_000, 500 if a.small else 2000, spectral=(name == "c5s"))
        r["n_gpus"] = world
        if rank == 0:
            print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
