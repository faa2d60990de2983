// K6: landmark operator.  Segmented column aggregation of a CSR kernel by cluster label
// (transitions = row-normalised  K . C), column sums, and the dense landmark operator
// landmark_op = rownorm(C^T K) . rownorm(K C).
//
// Replaces LandmarkGraph._landmarks_to_data / build_landmark_op / extend_to_data
// (reference graphtools/graphs.py:1169-1182, :1232-1246, :1272-1288): the per-cluster boolean row
// slicing + vstack, sklearn normalize and the scipy sparse product.
//
// K is symmetric on this path (the host checks), so pnm[j, l] = sum_{i in cluster l} K[i, j] is the
// aggregation of ROW j of K by the label of the column: one warp per row, no transpose.  Within a label the
// entries are added in column order -- the order in which the reference's kernel[clusters == l, :].sum(axis=0)
// adds the rows of a cluster.
//
// Every cross-row reduction (column sums of pnm, the L x L operator) is accumulated in FIXED POINT with 64-bit
// integer atomics -- a 96-bit value split over two words.  Integer addition is associative, so the result does not
// depend on the order in which the atomics land: landmark_op is bit-reproducible from run to run (and would be
// across any partition of the rows), and the sums are exact to 2^-94 (2^-72 for the column sums) before the single
// final rounding to float64 -- more accurate than a float64 running sum.
#include "common.cuh"
#include "gtb200.h"

namespace {

constexpr int AGG_WARPS = 4, AGG_CAP = 256, AGG_LONG_THREADS = 256;
typedef unsigned long long u64;

// value v >= 0 with v * 2^HI_BITS < 2^63  ->  (floor(v 2^HI), floor(frac 2^32)): hi words add up to < 2^63 as long as the
// total stays below 2^(63 - HI_BITS); lo words hold 32 bits each, so 2^32 of them fit a 64-bit sum.
template <int HI_BITS>
__device__ __forceinline__ void fx_split(double v, u64& hi, u64& lo) {
  const double s = ldexp(v, HI_BITS);
  const double f = floor(s);
  hi = (u64)f;
  lo = (u64)ldexp(s - f, 32);
}
template <int HI_BITS>
__device__ __forceinline__ double fx_join(u64 hi, u64 lo) {
  // carry the overflow of the low word into the high one, then round once
  hi += lo >> 32;
  lo &= 0xffffffffull;
  return ldexp((double)hi, -HI_BITS) + ldexp((double)lo, -(HI_BITS + 32));
}
constexpr int COLSUM_HI = 40;   // column sums of pnm stay below 2^23 (<= number of samples)
constexpr int OP_HI = 62;       // entries of landmark_op (a row-stochastic product) stay <= 1

// key = (label << 32) | position-in-row: sorting by it groups the labels and keeps column order inside a label
struct RowAgg {
  const int64_t* indptr; const int32_t* idx; const double* val; int64_t n; const int32_t* label;
  int32_t* cnt; const int64_t* outptr; int32_t* out_idx; double* out_raw; double* out_norm; u64* colsum_fx;
  int32_t* has_long;
};

template <bool FILL>
__global__ void __launch_bounds__(AGG_WARPS * 32) cluster_aggregate_kernel(RowAgg p) {
  __shared__ u64 ks[AGG_WARPS][AGG_CAP];
  __shared__ double vs[AGG_WARPS][AGG_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * AGG_WARPS + warp;
  if (row >= p.n) return;
  const int64_t p0 = p.indptr[row];
  const int64_t L64 = p.indptr[row + 1] - p0;
  if (L64 > AGG_CAP) {                        // left to cluster_aggregate_long_kernel
    if (lane == 0) *p.has_long = 1;
    return;
  }
  const int L = (int)L64;
  if (L == 0) {
    if (!FILL && lane == 0) p.cnt[row] = 0;
    return;
  }
  u64* k = ks[warp];
  double* v = vs[warp];
  if (L <= 32) {
    // one entry per lane; its place in (label, column) order by counting -- no network
    int32_t lab = 0x7fffffff;
    double w = 0.0;
    if (lane < L) { lab = p.label[p.idx[p0 + lane]]; w = p.val[p0 + lane]; }
    int rank = 0;
    for (int t = 0; t < L; ++t) {
      const int32_t lt = __shfl_sync(0xffffffffu, lab, t);
      rank += (lt < lab) | ((lt == lab) & (t < lane));
    }
    if (lane < L) { k[rank] = (u64)(uint32_t)lab; v[rank] = w; }
    __syncwarp();
  } else {
    int np2 = 64;
    while (np2 < L) np2 <<= 1;
    for (int t = lane; t < np2; t += 32) {
      if (t < L) { k[t] = ((u64)(uint32_t)p.label[p.idx[p0 + t]] << 32) | (u64)t; v[t] = p.val[p0 + t]; }
      else { k[t] = ~0ull; v[t] = 0.0; }
    }
    __syncwarp();
    auto sync = [] { __syncwarp(); };
    GTB_BITONIC_SORT(k, v, np2, lane, 32, sync, u64, double);
    for (int t = lane; t < L; t += 32) k[t] >>= 32;
    __syncwarp();
  }
  // k[0..L) = labels ascending, v = values in column order inside each label.  Segment heads -> one output each.
  int base = 0;
  double rs_part = 0.0;                        // this lane's share of the row's L1 norm
  const int64_t o0 = FILL ? p.outptr[row] : 0;
  for (int t0 = 0; t0 < L; t0 += 32) {
    const int t = t0 + lane;
    const bool head = (t < L) && (t == 0 || k[t] != k[t - 1]);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (FILL && head) {
      double s = 0.0;
      for (int u = t; u < L && k[u] == k[t]; ++u) s += v[u];       // sequential, column order
      const int o = base + __popc(hm & ((1u << lane) - 1u));
      p.out_idx[o0 + o] = (int32_t)k[t];
      p.out_raw[o0 + o] = s;
      rs_part += fabs(s);
    }
    base += __popc(hm);
  }
  if (!FILL) {
    if (lane == 0) p.cnt[row] = base;
    return;
  }
  // row L1 norm: fixed-order reduction of the per-head sums (heads were assigned to lanes by position)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) rs_part += __shfl_xor_sync(0xffffffffu, rs_part, off);
  __syncwarp();
  for (int t = lane; t < base; t += 32) {
    const double s = p.out_raw[o0 + t];
    if (p.out_norm) p.out_norm[o0 + t] = (rs_part != 0.0) ? s / rs_part : s;
    if (p.colsum_fx) {
      u64 hi, lo;
      fx_split<COLSUM_HI>(fabs(s), hi, lo);
      const int32_t l = p.out_idx[o0 + t];
      atomicAdd(p.colsum_fx + 2 * (int64_t)l, hi);
      atomicAdd(p.colsum_fx + 2 * (int64_t)l + 1, lo);
    }
  }
}

// Rows longer than AGG_CAP (every row of a dense kernel handed over as CSR; hub rows): one block per row walks the
// entries once and accumulates per label into a dense fixed-point table of n_label slots (shared memory when it
// fits, else this block's slice of `ws`), then emits the occupied slots in label order: O(len + n_label) per row.
__global__ void __launch_bounds__(AGG_LONG_THREADS) cluster_aggregate_long_kernel(RowAgg p, int n_label, int fill,
                                                                                 u64* __restrict__ ws,
                                                                                 int table_in_smem) {
  extern __shared__ __align__(16) unsigned char agg_smem[];
  __shared__ unsigned int long_mask[AGG_LONG_THREADS / 32];
  __shared__ int scan_s[AGG_LONG_THREADS / 32];
  __shared__ int base_s;
  __shared__ double rs_s[AGG_LONG_THREADS / 32];
  if (*p.has_long == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* table = table_in_smem ? reinterpret_cast<u64*>(agg_smem) : (ws + (size_t)blockIdx.x * 3 * (size_t)n_label);
  u64* thi = table;
  u64* tlo = table + n_label;
  u64* tocc = table + 2 * (size_t)n_label;           // number of entries that hit the slot (occupancy)
  for (int64_t b0 = (int64_t)blockIdx.x * AGG_LONG_THREADS; b0 < p.n; b0 += (int64_t)gridDim.x * AGG_LONG_THREADS) {
    const int64_t r = b0 + tid;
    const bool is_long = (r < p.n) && (p.indptr[r + 1] - p.indptr[r] > AGG_CAP);
    const unsigned m = __ballot_sync(0xffffffffu, is_long);
    if (lane == 0) long_mask[warp] = m;
    __syncthreads();
    for (int w = 0; w < AGG_LONG_THREADS / 32; ++w) {
      unsigned mm = long_mask[w];
      while (mm) {
        const int bit = __ffs(mm) - 1;
        mm &= mm - 1;
        const int64_t row = b0 + w * 32 + bit;
        const int64_t p0 = p.indptr[row], p1 = p.indptr[row + 1];
        for (int l = tid; l < n_label; l += AGG_LONG_THREADS) { thi[l] = 0; tlo[l] = 0; tocc[l] = 0; }
        __syncthreads();
        for (int64_t e = p0 + tid; e < p1; e += AGG_LONG_THREADS) {
          const int32_t l = p.label[p.idx[e]];
          u64 hi, lo;
          fx_split<COLSUM_HI>(fabs(p.val[e]), hi, lo);
          atomicAdd(thi + l, hi);
          atomicAdd(tlo + l, lo);
          atomicAdd(tocc + l, 1ull);
        }
        __syncthreads();
        // occupied slots in label order: block-wide exclusive count, chunk by chunk
        if (tid == 0) base_s = 0;
        double rs = 0.0;
        __syncthreads();
        const int64_t o0 = fill ? p.outptr[row] : 0;
        for (int l0 = 0; l0 < n_label; l0 += AGG_LONG_THREADS) {
          const int l = l0 + tid;
          const bool occ = (l < n_label) && (tocc[l] != 0);
          const unsigned om = __ballot_sync(0xffffffffu, occ);
          if (lane == 0) scan_s[warp] = __popc(om);
          __syncthreads();
          int before = base_s;
          for (int q = 0; q < warp; ++q) before += scan_s[q];
          if (fill && occ) {
            const int o = before + __popc(om & ((1u << lane) - 1u));
            const double s = fx_join<COLSUM_HI>(thi[l], tlo[l]);
            p.out_idx[o0 + o] = l;
            p.out_raw[o0 + o] = s;
            rs += s;
          }
          __syncthreads();
          if (tid == 0) { int tot = 0; for (int q = 0; q < AGG_LONG_THREADS / 32; ++q) tot += scan_s[q]; base_s += tot; }
          __syncthreads();
        }
        const int heads = base_s;
        if (!fill) {
          if (tid == 0) p.cnt[row] = heads;
        } else {
          // row L1 norm, fixed-order reduction
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
          if (lane == 0) rs_s[warp] = rs;
          __syncthreads();
          double tot = 0.0;
          for (int q = 0; q < AGG_LONG_THREADS / 32; ++q) tot += rs_s[q];
          for (int t = tid; t < heads; t += AGG_LONG_THREADS) {
            const double s = p.out_raw[o0 + t];
            if (p.out_norm) p.out_norm[o0 + t] = (tot != 0.0) ? s / tot : s;
            if (p.colsum_fx) {
              u64 hi, lo;
              fx_split<COLSUM_HI>(s, hi, lo);
              const int32_t l = p.out_idx[o0 + t];
              atomicAdd(p.colsum_fx + 2 * (int64_t)l, hi);
              atomicAdd(p.colsum_fx + 2 * (int64_t)l + 1, lo);
            }
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

__global__ void colsum_finalize_kernel(const u64* __restrict__ fx, int n_label, double* __restrict__ colsum) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_label) colsum[l] = fx_join<COLSUM_HI>(fx[2 * l], fx[2 * l + 1]);
}

// op[l][m] += (raw[j,l] / colsum[l]) * norm[j,m] over the non-zeros of row j, accumulated in fixed point
__global__ void __launch_bounds__(256) landmark_op_kernel(const int64_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ lab,
                                                          const double* __restrict__ raw,
                                                          const double* __restrict__ nrm,
                                                          const double* __restrict__ colsum, int64_t n, int L,
                                                          u64* __restrict__ op_fx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int64_t p0 = ptr[row];
  const int len = (int)(ptr[row + 1] - p0);
  const int64_t npair = (int64_t)len * len;
  for (int64_t q = lane; q < npair; q += 32) {
    const int a = (int)(q / len), b = (int)(q - (int64_t)a * len);
    const int la = lab[p0 + a];
    const double cs = colsum[la];
    const double left = (cs != 0.0) ? raw[p0 + a] / cs : raw[p0 + a];
    u64 hi, lo;
    fx_split<OP_HI>(fabs(left * nrm[p0 + b]), hi, lo);
    u64* dst = op_fx + 2 * ((int64_t)la * L + lab[p0 + b]);
    atomicAdd(dst, hi);
    atomicAdd(dst + 1, lo);
  }
}

__global__ void landmark_op_finalize_kernel(const u64* __restrict__ fx, int64_t n2, double* __restrict__ op) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) op[i] = fx_join<OP_HI>(fx[2 * i], fx[2 * i + 1]);
}

constexpr int AGG_SMEM_LABELS = 4096;   // 3 x 8 bytes per label -> 96 KB of dynamic shared memory

int launch_long(const RowAgg& p, int n_label, int fill, u64* ws, cudaStream_t st) {
  const bool in_smem = n_label <= AGG_SMEM_LABELS;
  const size_t smem = in_smem ? (size_t)3 * n_label * sizeof(u64) : 0;
  if (in_smem) {
    static size_t attr = 0;
    if (smem > attr) {
      GTB_CUDA(cudaFuncSetAttribute(cluster_aggregate_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)((size_t)3 * AGG_SMEM_LABELS * sizeof(u64))));
      attr = (size_t)3 * AGG_SMEM_LABELS * sizeof(u64);
    }
  }
  const int64_t want = gtb_cdiv(p.n, AGG_LONG_THREADS);
  const unsigned grid = (unsigned)(want < 296 ? want : 296);
  cluster_aggregate_long_kernel<<<grid, AGG_LONG_THREADS, smem, st>>>(p, n_label, fill, ws, in_smem ? 1 : 0);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

}  // namespace

// scratch (in 8-byte words) of the aggregate calls: [0] = long-row flag, then the per-block dense tables of the
// long-row kernel when n_label does not fit shared memory
extern "C" int64_t gtb_cluster_aggregate_ws_elems(int n_label) {
  return 1 + (n_label > AGG_SMEM_LABELS ? (int64_t)296 * 3 * n_label : 0);
}

extern "C" int gtb_cluster_aggregate_count(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n,
                                           const int32_t* label, int n_label, int32_t* cnt, void* ws,
                                           void* stream) {
  GTB_CHECK_ARG(n > 0 && n_label > 0, "empty matrix");
  cudaStream_t st = (cudaStream_t)stream;
  RowAgg p{};
  p.indptr = indptr; p.idx = idx; p.val = val; p.n = n; p.label = label; p.cnt = cnt;
  p.has_long = reinterpret_cast<int32_t*>(ws);
  GTB_CUDA(cudaMemsetAsync(ws, 0, 8, st));
  cluster_aggregate_kernel<false><<<(unsigned)gtb_cdiv(n, AGG_WARPS), AGG_WARPS * 32, 0, st>>>(p);
  GTB_CHECK_LAUNCH();
  return launch_long(p, n_label, 0, reinterpret_cast<u64*>(ws) + 1, st);
}

extern "C" int gtb_cluster_aggregate_fill(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n,
                                          const int32_t* label, const int64_t* outptr, int32_t* out_idx,
                                          double* out_raw, double* out_norm, double* colsum, void* colsum_fx,
                                          int n_label, void* ws, void* stream) {
  GTB_CHECK_ARG(n > 0 && n_label > 0, "empty matrix");
  GTB_CHECK_ARG(colsum == nullptr || colsum_fx != nullptr, "colsum needs its fixed-point scratch (2 x n_label words)");
  cudaStream_t st = (cudaStream_t)stream;
  RowAgg p{};
  p.indptr = indptr; p.idx = idx; p.val = val; p.n = n; p.label = label; p.outptr = outptr; p.out_idx = out_idx;
  p.out_raw = out_raw; p.out_norm = out_norm; p.colsum_fx = colsum ? reinterpret_cast<u64*>(colsum_fx) : nullptr;
  p.has_long = reinterpret_cast<int32_t*>(ws);
  GTB_CUDA(cudaMemsetAsync(ws, 0, 8, st));
  if (colsum) GTB_CUDA(cudaMemsetAsync(colsum_fx, 0, sizeof(u64) * 2 * (size_t)n_label, st));
  cluster_aggregate_kernel<true><<<(unsigned)gtb_cdiv(n, AGG_WARPS), AGG_WARPS * 32, 0, st>>>(p);
  GTB_CHECK_LAUNCH();
  int rc = launch_long(p, n_label, 1, reinterpret_cast<u64*>(ws) + 1, st);
  if (rc) return rc;
  if (colsum) {
    colsum_finalize_kernel<<<(unsigned)gtb_cdiv(n_label, 256), 256, 0, st>>>(reinterpret_cast<const u64*>(colsum_fx),
                                                                            n_label, colsum);
    GTB_CHECK_LAUNCH();
  }
  return GTB_OK;
}

extern "C" int gtb_landmark_op(const int64_t* ptr, const int32_t* lab, const double* raw, const double* nrm,
                               const double* colsum, int64_t n, int L, double* op, void* op_fx, void* stream) {
  GTB_CHECK_ARG(n > 0 && L > 0, "empty input");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(op_fx, 0, sizeof(u64) * 2 * (size_t)L * L, st));
  landmark_op_kernel<<<(unsigned)gtb_cdiv(n * 32, 256), 256, 0, st>>>(ptr, lab, raw, nrm, colsum, n, L,
                                                                      reinterpret_cast<u64*>(op_fx));
  GTB_CHECK_LAUNCH();
  const int64_t n2 = (int64_t)L * L;
  landmark_op_finalize_kernel<<<(unsigned)gtb_cdiv(n2, 256), 256, 0, st>>>(reinterpret_cast<const u64*>(op_fx), n2, op);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
