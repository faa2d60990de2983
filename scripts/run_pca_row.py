"""f4 row: device PCA front-end vs scikit-learn on the host, same seed (one JSON line)."""
import json, os, sys, time, warnings
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphtools_b200 import pca, synth
warnings.simplefilter("ignore")
n, d, k = 100_000, 1000, 100
X, _ = synth.gaussian_mixture(n, d, n_clusters=20, intrinsic_dim=30, seed=1)
X = X.astype(np.float64)
pca.fit_transform_dense(X[:2000], 20, 0)                      # warm-up (cuBLAS / cuSOLVER handles)
torch.cuda.synchronize()
t0 = time.perf_counter()
op, Z = pca.fit_transform_dense(X, k, 42)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
Xd = torch.from_numpy(X).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
op2, Z2 = pca.fit_transform_dense(Xd, k, 42)
torch.cuda.synchronize()
t_dev_res = time.perf_counter() - t0
from sklearn.decomposition import PCA
t0 = time.perf_counter()
ref = PCA(k, svd_solver="randomized", random_state=42).fit(X)
Zr = ref.transform(X)
t_host = time.perf_counter() - t0
err = float(np.abs(Z.cpu().numpy() - Zr).max() / np.abs(Zr).max())
print(json.dumps({"row": "f4 PCA front-end (randomized SVD, n_pca=100) on dense float64 %dx%d" % (n, d),
                  "device_s_incl_h2d": t_dev, "device_s_input_resident": t_dev_res, "host_sklearn_s": t_host,
                  "host_threads": os.cpu_count(), "max_abs_diff_over_max_abs": err,
                  "singular_values_rel_diff": float(np.abs(op.singular_values_ / ref.singular_values_ - 1).max())}))
