"""CUDA-event breakdown of the row-sharded build at N ranks (run under torchrun): where the step goes outside the
search kernel -- the shard upload + NCCL all-gather of the reference set, the bucketing kernels, the two all-to-all
calls of the edge exchange, the receive-side transpose, the merge, and the shared-memory host assembly.  Each stage
is timed with events on the launching stream; the line printed by rank 0 carries the MAX over ranks per stage.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/exchange_breakdown.py [--size 1000000] [--reps 3]
"""
import argparse
import json
import os
import sys
import time
import warnings

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphtools_b200 as gt
from graphtools_b200 import _engine as E, distributed as gd, pipeline, synth

warnings.simplefilter("ignore")
STAGES = {}


def timed(mod, name, label):
    fn = getattr(mod, name)

    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        STAGES.setdefault(label, []).append((e0, e1))
        return out
    setattr(mod, name, wrapper)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    timed(gd, "upload_sharded", "reference set: shard H2D + NCCL all-gather (collective 1)")
    timed(gd, "cuda_bucket_edges", "edge exchange: bucket by owner (route_count / scan / route_fill)")
    timed(dist, "all_to_all_single", "edge exchange: NCCL all-to-all (split sizes, then packed records; collective 2)")
    timed(pipeline, "merge_with_transpose", "merge + normalise per shard (sym_merge)")
    X, _ = synth.gaussian_mixture(a.n, 100, n_clusters=50, intrinsic_dim=10, seed=3)
    Xh = torch.from_numpy(X).pin_memory()
    rows = []
    for rep in range(a.reps + 2):          # two warm-up builds: the pooled host segments exist and alternate afterwards
        STAGES.clear()
        E.timing = {}
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ev0.record()
        G = gt.Graph(Xh, knn=5, decay=40, thresh=1e-4, verbose=0)
        G._ensure_built()
        ev1.record()
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        K, P = G.kernel, G.diff_op                    # shared-memory host assembly (every rank writes its shard)
        dist.barrier()
        t_all = time.perf_counter() - t0
        tm = E.timings_ms()
        E.timing = None
        st = {k: sum(x.elapsed_time(y) for x, y in v) for k, v in STAGES.items()}
        st["device build (events, this rank)"] = ev0.elapsed_time(ev1)
        st["host wall: build"] = 1e3 * t_build
        st["host wall: build + K/P assembled on the host (all ranks)"] = 1e3 * t_all
        st["host assembly of K / P (shared memory, D2H per rank)"] = 1e3 * (t_all - t_build)
        for k in ("gtb_knn_topk_tc_seeded", "gtb_knn_seed_tc", "gtb_knn_topk_tc", "gtb_refine_topk",
                  "gtb_records_count", "gtb_records_scatter", "gtb_rec_sort_rows", "gtb_sym_merge_count",
                  "gtb_sym_merge_fill", "gtb_route_count", "gtb_route_fill", "gtb_prepare_operand_tc"):
            if k in tm:
                st["kernel " + k] = tm[k][1]
        if rep >= 2:
            rows.append(st)
    keys = sorted(set().union(*[r.keys() for r in rows]))
    mine = torch.tensor([[r.get(k, 0.0) for k in keys] for r in rows], dtype=torch.float64, device="cuda").mean(0)
    mx = mine.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "n": a.n, "reps": a.reps, "max_over_ranks_ms": dict(zip(keys, [round(v, 3) for v in mx.tolist()])),
                          "rank0_ms": dict(zip(keys, [round(v, 3) for v in mine.tolist()])), "nnz": int(K.nnz)}), flush=True)
    dist.destroy_process_group()
