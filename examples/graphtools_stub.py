"""The ctypes stub of INTEGRATION.md section 2, verbatim (kept executable so the document cannot rot:
tests/test_integration_stub_gpu.py runs it against the oracle)."""
# graphtools/_gtb200.py
import ctypes, numpy as np, torch
from scipy import sparse

import os
_L = ctypes.CDLL(os.environ.get("GTB200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "graphtools_b200", "libgtb200.so")))
_L.gtb_last_error.restype = ctypes.c_char_p
_L.gtb_col_mean_ws_doubles.restype = ctypes.c_int64
_L.gtb_scan_ws_elems.restype = ctypes.c_int64
_L.gtb_tc_scratch_bytes.restype = ctypes.c_int64
P = ctypes.c_void_p; I64 = ctypes.c_int64; I = ctypes.c_int; D = ctypes.c_double; F = ctypes.c_float

def _call(name, *args):
    rc = getattr(_L, name)(*args, P(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(_L.gtb_last_error().decode())

def p(t):                       # tensor -> void*
    return P(t.data_ptr()) if t is not None else P(0)

def knn_alpha_decay_kernel(X, Y, knn, decay, thresh, bandwidth_scale=1.0):
    """CSR [n_y, n] of exp(-(d/bw)^decay) >= thresh, bw = distance to the knn-th neighbour
    (graphs.py:886-897); X, Y float32-exact host arrays, d + 1 <= 104 (3xTF32 operands, dtype 0)."""
    dev = torch.device("cuda")
    Xd = torch.as_tensor(np.ascontiguousarray(X, np.float32), device=dev)
    Yd = Xd if Y is X else torch.as_tensor(np.ascontiguousarray(Y, np.float32), device=dev)
    n, d = Xd.shape; ny = Yd.shape[0]
    npad = lambda m: (m + 127) // 128 * 128
    Kp = (d + 1 + 7) // 8 * 8
    e = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=dev)
    mean = e(d); ws = e(_L.gtb_col_mean_ws_doubles(I(d)), dt=torch.float64)
    _call("gtb_col_mean", p(Xd), I64(n), I(d), p(ws), p(mean))
    def operand(A, role):
        m = A.shape[0]; hi, lo, n2, mx = e(npad(m), Kp), e(npad(m), Kp), e(npad(m)), e(1)
        _call("gtb_prepare_operand_tc", p(A), I64(m), I(d), p(mean), I(role), p(hi), p(lo), I64(npad(m)), I(Kp),
              I(0), F(1.0), p(n2), p(mx))        # dtype 0 = tf32 hi/lo pairs in float32 (1 = bfloat16 pairs, Kp % 16 == 0)
        return hi, lo, n2, mx
    q_hi, q_lo, q_n2, _ = operand(Yd, 0)
    r_hi, r_lo, _, r_max = operand(Xd, 1)
    cand = e(ny, 64, dt=torch.int32); tau = e(ny, 2)
    scratch = e(_L.gtb_tc_scratch_bytes(I64(npad(ny))), dt=torch.uint8)
    _call("gtb_knn_topk_tc", p(q_hi), p(q_lo), p(q_n2), I64(ny), I64(npad(ny)), p(r_hi), p(r_lo), I64(n),
          I64(npad(n)), I(Kp), I(0), I(32), I(2), I(1), p(cand), p(scratch), p(tau), P(0))
    st_idx = e(ny, 64, dt=torch.int32); st_val = e(ny, 64, dt=torch.float64)
    n_keep, status, nzero = (e(ny, dt=torch.int32) for _ in range(3))
    bw = e(ny, dt=torch.float64); lim2 = e(ny)
    eps_rel = 4.0 * (d + 16) * 2.0 ** -24
    _call("gtb_refine_topk", p(Yd), I64(ny), p(Xd), I(d), I(0), p(cand), I(64), I(64), p(tau), I(2), p(q_n2),
          F(float(r_max.item())), D(eps_rel), I(knn), I64(0), D(decay), D(thresh), P(0), I(0), D(bandwidth_scale),
          p(st_idx), p(st_val), p(n_keep), p(bw), p(lim2), p(status), p(nzero))
    if int((status != 1).sum()):      # uncertified rows: radius pass, see graphtools_b200/pipeline.py:knn_kernel
        raise NotImplementedError("call gtb_knn_radius_tc + gtb_refine_ball for these rows")
    indptr = e(ny + 1, dt=torch.int64); sws = e(_L.gtb_scan_ws_elems(I64(ny)), dt=torch.int64)
    _call("gtb_exclusive_scan", p(n_keep), I64(ny), p(indptr), p(sws))
    nnz = int(indptr[-1])
    idx = e(nnz, dt=torch.int32); val = e(nnz, dt=torch.float64)
    _call("gtb_csr_gather", p(st_idx), p(st_val), p(n_keep), p(status), p(indptr), I64(ny), I(64), P(0), I64(0),
          P(0), P(0), P(0), P(0), p(idx), p(val))
    return sparse.csr_matrix((val.cpu().numpy(), idx.cpu().numpy(), indptr.cpu().numpy().astype(np.int32)),
                             shape=(ny, n))
