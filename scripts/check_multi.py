"""Multi-GPU end-to-end check (run under torchrun, N >= 2): the row-sharded build from HOST data -- shard upload +
NCCL all-gather of the reference set, kernel-bucketed edge exchange, per-shard merge, shared-memory host assembly --
must give K, P and the degree vector bit-identical to the single-GPU build of the same data on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_multi.py [--size 200000]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphtools_b200 as gt
from graphtools_b200 import synth

warnings.simplefilter("ignore")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=200_000)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    world = dist.get_world_size()
    out = {"n_gpus": world, "n": a.n}
    for symm, theta in (("+", None), ("mnn", 0.4), ("*", None)):
        X, _ = synth.gaussian_mixture(a.n, 100, n_clusters=20, intrinsic_dim=10, seed=4)
        Xh = torch.from_numpy(X).pin_memory()
        kw = dict(knn=5, decay=40, thresh=1e-4, kernel_symm=symm, theta=theta, verbose=0)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        G = gt.Graph(Xh, **kw)
        K, P, deg = G.kernel, G.diff_op, G.kernel_degree
        dist.barrier()
        t_sharded = time.perf_counter() - t0
        assert "_dev_shard" in G.__dict__ and "_dev_kernel" not in G.__dict__, "the build was not sharded"
        os.environ["GTB_DISTRIBUTED"] = "0"
        try:
            t0 = time.perf_counter()
            G1 = gt.Graph(Xh, **kw)
            K1, P1, deg1 = G1.kernel, G1.diff_op, G1.kernel_degree
            torch.cuda.synchronize()
            t_single = time.perf_counter() - t0
        finally:
            os.environ.pop("GTB_DISTRIBUTED")
        same = (np.array_equal(K.indptr, K1.indptr) and np.array_equal(K.indices, K1.indices)
                and np.array_equal(K.data, K1.data) and np.array_equal(P.data, P1.data)
                and np.array_equal(np.asarray(deg), np.asarray(deg1)))
        # the device-side gather used by landmark / MNN consumers
        Kd = G._dev_kernel
        same_dev = (np.array_equal(Kd.indices.cpu().numpy(), K1.indices) and np.array_equal(Kd.data.cpu().numpy(), K1.data)
                    and np.array_equal(G._dev_P.cpu().numpy(), P1.data))
        flag = torch.tensor([int(same and same_dev)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["symm_%s" % symm] = {"bit_identical_on_every_rank": bool(flag.item()), "nnz": int(K.nnz),
                                 "sharded_e2e_s": t_sharded, "single_gpu_e2e_s": t_single,
                                 "writable_rank0_only": bool(K.data.flags.writeable == (rank == 0))}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
