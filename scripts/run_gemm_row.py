"""Measurement of the landmark diffusion chain to the bar of the hot path: device time of one
float64-faithful product and of landmark_op^t (CUDA events), issued int8 tensor work against the nominal int8 peak,
float64-equivalent rate, numpy's time for the same product on the host beside it, and the parity figure."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphtools_b200 import dense, pipeline


def ev_time(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=2000)
    ap.add_argument("--t", type=int, default=100)
    a = ap.parse_args()
    L = a.L
    rng = np.random.default_rng(0)
    P = rng.random((L, L)) ** 12
    P /= P.sum(1, keepdims=True)
    Pd = pipeline.to_device(P)
    S = dense.GEMM_SLICES
    Lp = (L + 127) // 128 * 128
    pairs = S * (S + 1) // 2
    t0 = time.perf_counter()
    ref2 = P @ P
    t_np = time.perf_counter() - t0
    for rb in (32, 64, 128):
        da, db = dense.slice_operand(Pd, False), dense.slice_operand(Pd, True)
        ms_k, C = ev_time(lambda: dense.gemm_digits(da, db, L, L, row_bytes=rb))
        ms_all, _ = ev_time(lambda: dense.gemm_f64(Pd, Pd, row_bytes=rb))
        err = float(np.abs(C.cpu().numpy() - ref2).max() / np.abs(ref2).max())
        issued = 2.0 * pairs * Lp ** 3
        print(json.dumps({"row": "f3 float64-faithful product on int8 tensor cores (gtb_gemm_i8)", "L": L, "row_bytes": rb,
                          "slices": S, "digit_pairs": pairs, "kernel_ms": ms_k, "with_slicing_ms": ms_all,
                          "f64_equivalent_TFLOPs": 2.0 * L ** 3 / ms_k / 1e9, "issued_int8_TOPs": issued / ms_k / 1e9,
                          "nominal_int8_peak_TOPs": 4500.0, "frac_of_nominal_int8_peak": issued / ms_k / 1e9 / 4500.0,
                          "numpy_host_ms": 1e3 * t_np, "max_err_over_max": err}), flush=True)
    t0 = time.perf_counter()
    ref = np.linalg.matrix_power(P, a.t)
    t_np = time.perf_counter() - t0
    ms, out = ev_time(lambda: dense.matrix_power(Pd, a.t), reps=5)
    got = out.cpu().numpy()
    print(json.dumps({"row": "f3 landmark_op^t (dense.matrix_power, numpy's schedule)", "L": L, "t": a.t, "device_ms": ms,
                      "numpy_host_ms": 1e3 * t_np, "max_rel_err": float((np.abs(got - ref) / ref).max()),
                      "rows_sum_to_one": float(np.abs(got.sum(1) - 1).max())}), flush=True)


if __name__ == "__main__":
    main()
