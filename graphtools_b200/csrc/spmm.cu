// CSR (float64) x dense (float64, row-major) product on the device: out = A . B.
//
// Replaces scipy's csr_matvecs behind `transitions.dot(transform)` (reference graphtools/base.py:1229),
// the repeated `diff_op.dot(X)` of the callers (MAGIC-style diffusion) and the two sparse products per power
// iteration of sklearn randomized_svd(diff_aff) (graphs.py:1217).  HBM / gather bound: every stored entry pulls
// one f-wide row of B (8f bytes, coalesced), so algorithmic bytes = nnz (12 + 8f) + 8 n_rows f.
//
// Arithmetic contract: each output element is accumulated in stored (column-sorted) order with a separate
// multiply and add (no FMA contraction) -- the order and rounding of scipy's axpy loop, so results are
// bit-identical to `A.dot(B)` on the host.
//
// Layout: one warp per output row; lane l owns columns {2l, 2l+1} of every 64-column chunk (double2 loads:
// 512 contiguous bytes per warp and chunk).  Entries are consumed four at a time so that four independent row
// gathers are in flight per warp.
#include "common.cuh"
#include "gtb200.h"

namespace {

constexpr int SPMM_WARPS = 8;
constexpr int SPMM_UNROLL = 4;

// NCH = 64-column chunks held in registers by one warp pass (f <= 64 * NCH handled in one sweep over the row)
template <int NCH>
__global__ void __launch_bounds__(SPMM_WARPS * 32)
spmm_csr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                int64_t n_rows, const double* __restrict__ B, int64_t ldb, int f, int col0, double* __restrict__ out,
                int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t p0 = indptr[row], p1 = indptr[row + 1];
  double ax[NCH], ay[NCH];
  bool okx[NCH], oky[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    ax[c] = 0.0; ay[c] = 0.0;
    const int k = col0 + c * 64 + 2 * lane;
    okx[c] = k < f; oky[c] = k + 1 < f;
  }
  const bool vec = ((ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  for (int64_t p = p0; p < p1; p += SPMM_UNROLL) {
    double a[SPMM_UNROLL];
    const double* src[SPMM_UNROLL];
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; ++u) {
      const bool in = p + u < p1;
      a[u] = in ? val[p + u] : 0.0;
      src[u] = B + (int64_t)(in ? idx[p + u] : 0) * ldb + col0 + 2 * lane;
    }
    double bx[SPMM_UNROLL][NCH], by[SPMM_UNROLL][NCH];
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; ++u) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        bx[u][c] = 0.0; by[u][c] = 0.0;
        if (p + u < p1) {
          if (vec && oky[c]) {
            const double2 t = *reinterpret_cast<const double2*>(src[u] + c * 64);
            bx[u][c] = t.x; by[u][c] = t.y;
          } else {
            if (okx[c]) bx[u][c] = src[u][c * 64];
            if (oky[c]) by[u][c] = src[u][c * 64 + 1];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; ++u) {
      if (p + u < p1) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          ax[c] = __dadd_rn(ax[c], __dmul_rn(a[u], bx[u][c]));
          ay[c] = __dadd_rn(ay[c], __dmul_rn(a[u], by[u][c]));
        }
      }
    }
  }
  double* o = out + row * ldo + col0 + 2 * lane;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (okx[c]) o[c * 64] = ax[c];
    if (oky[c]) o[c * 64 + 1] = ay[c];
  }
}

// out[i][k] *= s[i]  /  out[i][k] = in[i][k] * s[i]   (row scaling used around the SpMM for D^-1/2 K D^-1/2 products)
__global__ void row_scale_kernel(const double* __restrict__ in, const double* __restrict__ s, int64_t n, int f,
                                 int power_neg_half, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * f) return;
  const int64_t r = e / f;
  double w = s[r];
  if (power_neg_half) w = 1.0 / sqrt(w);
  out[e] = in[e] * w;
}

}  // namespace

extern "C" int gtb_spmm_csr(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n_rows,
                            const double* B, int64_t ldb, int f, double* out, int64_t ldo, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && f > 0 && ldb >= f && ldo >= f, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)gtb_cdiv(n_rows, SPMM_WARPS);
  for (int col0 = 0; col0 < f; col0 += 128) {
    if (f - col0 > 64)
      spmm_csr_kernel<2><<<grid, SPMM_WARPS * 32, 0, st>>>(indptr, idx, val, n_rows, B, ldb, f, col0, out, ldo);
    else
      spmm_csr_kernel<1><<<grid, SPMM_WARPS * 32, 0, st>>>(indptr, idx, val, n_rows, B, ldb, f, col0, out, ldo);
    GTB_CHECK_LAUNCH();
  }
  return GTB_OK;
}

extern "C" int gtb_row_scale(const double* in, const double* s, int64_t n, int f, int power_neg_half, double* out,
                             void* stream) {
  GTB_CHECK_ARG(n > 0 && f > 0, "bad shape");
  row_scale_kernel<<<(unsigned)gtb_cdiv(n * f, 256), 256, 0, (cudaStream_t)stream>>>(in, s, n, f, power_neg_half, out);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
