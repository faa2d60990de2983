// K6: landmark operator.  Segmented column aggregation of a CSR kernel by cluster label
// (transitions = row-normalised  K . C), column sums, and the dense landmark operator
// landmark_op = rownorm(C^T K) . rownorm(K C).
//
// Replaces LandmarkGraph._landmarks_to_data / build_landmark_op / extend_to_data
// (reference graphtools/graphs.py:1169-1182, :1232-1246, :1272-1288): the per-cluster boolean row
// slicing + vstack, sklearn normalize and the scipy sparse product.
//
// K is symmetric on this path (the host checks), so pnm[j, l] = sum_{i in cluster l} K[i, j] is the
// aggregation of ROW j of K by the label of the column: one warp per row, no transpose.
#include "common.cuh"
#include "gtb200.h"

namespace {

constexpr int AGG_WARPS = 4, AGG_CAP = 64;

// Aggregates row `row` of (indptr, idx, val) by label[idx].  COUNT pass: cnt[row] = number of
// distinct labels.  FILL pass: writes (label, sum) sorted by label at outptr[row], the row-normalised
// value, and accumulates column sums.
template <bool FILL>
__global__ void __launch_bounds__(AGG_WARPS * 32) cluster_aggregate_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
    int64_t n, const int32_t* __restrict__ label, int32_t* __restrict__ cnt,
    const int64_t* __restrict__ outptr, int32_t* __restrict__ out_idx, double* __restrict__ out_raw,
    double* __restrict__ out_norm, double* __restrict__ colsum) {
  __shared__ int32_t ks[AGG_WARPS][AGG_CAP];
  __shared__ double vs[AGG_WARPS][AGG_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * AGG_WARPS + warp;
  if (row >= n) return;
  const int64_t p0 = indptr[row];
  const int64_t L = indptr[row + 1] - p0;
  auto sync = [] { __syncwarp(); };
  if (L <= AGG_CAP) {
    int32_t* k = ks[warp];
    double* v = vs[warp];
    int np2 = 2;
    while (np2 < L) np2 <<= 1;
    // key = label; ties keep column order through the payload-free stable trick: the columns are
    // already ascending, so encode position in the low bits of a 64-bit key is unnecessary --
    // summation order inside a label only changes the last ulp (tolerance is rtol 1e-5).
    for (int t = lane; t < np2; t += 32) {
      if (t < L) { k[t] = label[idx[p0 + t]]; v[t] = val[p0 + t]; }
      else { k[t] = 0x7fffffff; v[t] = 0.0; }
    }
    __syncwarp();
    GTB_BITONIC_SORT(k, v, np2, lane, 32, sync, int32_t, double);
    // segment heads
    int heads = 0;
    for (int t = lane; t < (int)L; t += 32) heads += (t == 0 || k[t] != k[t - 1]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, off);
    if (!FILL) {
      if (lane == 0) cnt[row] = heads;
      return;
    }
    // lane 0 walks the (short) sorted row sequentially: deterministic sums
    __shared__ double rowsum_s[AGG_WARPS];
    if (lane == 0) {
      const int64_t o0 = outptr[row];
      int o = 0;
      double rs = 0.0;
      int t = 0;
      while (t < (int)L) {
        int32_t lab = k[t];
        double s = 0.0;
        while (t < (int)L && k[t] == lab) { s += v[t]; ++t; }
        out_idx[o0 + o] = lab;
        out_raw[o0 + o] = s;
        rs += fabs(s);
        ++o;
      }
      rowsum_s[warp] = rs;
    }
    __syncwarp();
    const double rs = rowsum_s[warp];
    const int64_t o0 = outptr[row];
    for (int t = lane; t < heads; t += 32) {
      double s = out_raw[o0 + t];
      if (out_norm) out_norm[o0 + t] = (rs != 0.0) ? s / rs : s;
      if (colsum) atomicAdd(colsum + out_idx[o0 + t], fabs(s));
    }
  } else {
    // long rows: quadratic first-occurrence scheme straight from global memory
    int heads = 0;
    for (int64_t t = lane; t < L; t += 32) {
      const int32_t lab = label[idx[p0 + t]];
      bool first = true;
      for (int64_t u = 0; u < t; ++u) if (label[idx[p0 + u]] == lab) { first = false; break; }
      heads += first;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, off);
    if (!FILL) {
      if (lane == 0) cnt[row] = heads;
      return;
    }
    const int64_t o0 = outptr[row];
    double rs = 0.0;
    for (int64_t t = lane; t < L; t += 32) {
      const int32_t lab = label[idx[p0 + t]];
      bool first = true;
      for (int64_t u = 0; u < t; ++u) if (label[idx[p0 + u]] == lab) { first = false; break; }
      if (!first) continue;
      double s = 0.0;
      int rank = 0;
      for (int64_t u = 0; u < L; ++u) {
        const int32_t lu = label[idx[p0 + u]];
        if (lu == lab) s += val[p0 + u];
        else if (lu < lab) {
          bool f2 = true;  // count each smaller label once (at its first occurrence)
          for (int64_t w = 0; w < u; ++w) if (label[idx[p0 + w]] == lu) { f2 = false; break; }
          rank += f2;
        }
      }
      out_idx[o0 + rank] = lab;
      out_raw[o0 + rank] = s;
      rs += fabs(s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
    __syncwarp();
    for (int t = lane; t < heads; t += 32) {
      double s = out_raw[o0 + t];
      if (out_norm) out_norm[o0 + t] = (rs != 0.0) ? s / rs : s;
      if (colsum) atomicAdd(colsum + out_idx[o0 + t], fabs(s));
    }
  }
}

// op[l][m] += (raw[j,l] / colsum[l]) * norm[j,m] over the nonzeros of row j
__global__ void landmark_op_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ lab,
                                   const double* __restrict__ raw, const double* __restrict__ nrm,
                                   const double* __restrict__ colsum, int64_t n, int L, double* __restrict__ op) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int64_t p0 = ptr[row];
  const int len = (int)(ptr[row + 1] - p0);
  const int64_t npair = (int64_t)len * len;
  for (int64_t q = lane; q < npair; q += 32) {
    int a = (int)(q / len), b = (int)(q - (int64_t)a * len);
    int la = lab[p0 + a];
    double cs = colsum[la];
    double left = (cs != 0.0) ? raw[p0 + a] / cs : raw[p0 + a];
    atomicAdd(op + (int64_t)la * L + lab[p0 + b], left * nrm[p0 + b]);
  }
}

}  // namespace

extern "C" int gtb_cluster_aggregate_count(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n,
                                           const int32_t* label, int32_t* cnt, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  cluster_aggregate_kernel<false><<<(unsigned)gtb_cdiv(n, AGG_WARPS), AGG_WARPS * 32, 0, (cudaStream_t)stream>>>(
      indptr, idx, val, n, label, cnt, nullptr, nullptr, nullptr, nullptr, nullptr);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_cluster_aggregate_fill(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n,
                                          const int32_t* label, const int64_t* outptr, int32_t* out_idx,
                                          double* out_raw, double* out_norm, double* colsum, int n_label,
                                          void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  cudaStream_t st = (cudaStream_t)stream;
  if (colsum) GTB_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * n_label, st));
  cluster_aggregate_kernel<true><<<(unsigned)gtb_cdiv(n, AGG_WARPS), AGG_WARPS * 32, 0, st>>>(
      indptr, idx, val, n, label, nullptr, outptr, out_idx, out_raw, out_norm, colsum);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_landmark_op(const int64_t* ptr, const int32_t* lab, const double* raw, const double* nrm,
                               const double* colsum, int64_t n, int L, double* op, void* stream) {
  GTB_CHECK_ARG(n > 0 && L > 0, "empty input");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(op, 0, sizeof(double) * (size_t)L * L, st));
  landmark_op_kernel<<<(unsigned)gtb_cdiv(n * 32, 256), 256, 0, st>>>(ptr, lab, raw, nrm, colsum, n, L, op);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
