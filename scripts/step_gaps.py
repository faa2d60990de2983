"""Where the device step spends time outside the C-ABI kernels: torch.profiler over one Graph build (profiling aid)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphtools_b200 as gt
from graphtools_b200 import synth
from torch.profiler import profile, ProfilerActivity
X, _ = synth.gaussian_mixture(1_000_000, 100, n_clusters=50, intrinsic_dim=10, seed=3)
Xd = torch.from_numpy(X).cuda()
def step():
    G = gt.Graph(Xd, knn=5, decay=40, thresh=1e-4, verbose=0)
    return G._dev_kernel, G._dev_P
for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("wall one step %.1f ms" % ((time.perf_counter() - t0) * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=10, max_name_column_width=60))
