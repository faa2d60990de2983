import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import graphtools_b200 as gt
from graphtools_b200 import _engine as E, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
X, _ = synth.gaussian_mixture(n, 50, n_clusters=10, intrinsic_dim=10, seed=1)
Xd = torch.from_numpy(X).cuda()
for rep in range(2):
    E.timing = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    G = gt.Graph(Xd, graphtype="exact", knn=5, decay=40, thresh=1e-4, verbose=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    tm = E.timings_ms(); E.timing = None
    print("rep", rep, "wall %.1f ms" % ((t1 - t0) * 1e3), {k: round(v[1], 2) for k, v in tm.items()})
    del G
