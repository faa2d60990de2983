// K4: sparse symmetrisation (+, *, mnn), anisotropy and row-normalisation into the diffusion
// operator, plus the exclusive scan used to size every CSR in the engine.
//
// Replaces scipy's csr binops behind BaseGraph.symmetrize_kernel (reference graphtools/base.py:557-577,
// matrix.py:16-29), apply_anisotropy (base.py:579-592), sklearn normalize(K, "l1") (base.py:645) and
// kernel_degree (base.py:648-666).
//
// Symmetrisation never materialises K^T: every edge (i,j,w) binary-searches row j for column i
// (rows are column-sorted), which yields the reverse weight w' (0 when absent).  All three merge
// rules are symmetric functions s(w,w'), so an edge whose reverse is absent contributes the same
// value s to row j.  Pass 1 counts the new row lengths, pass 2 scatters (atomic cursor per row),
// pass 3 sorts each row by column and emits K, P = K / rowsum and the degree vector in one sweep.
#include "common.cuh"
#include "gtb200.h"

namespace {

// ------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                  int64_t* __restrict__ out,
                                                                  int64_t* __restrict__ blocksum) {
  __shared__ int64_t warp_tot[SCAN_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int64_t tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    tsum += v[i];
  }
  int64_t incl = tsum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int64_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
  int64_t run = wbase + incl - tsum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (tid == SCAN_THREADS - 1) blocksum[blockIdx.x] = wbase + incl;
}

// single block: exclusive scan of blocksum[0..nblk) in place; blocksum[nblk] = total
__global__ void scan_top_kernel(int64_t* __restrict__ blocksum, int64_t nblk) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t start = 0; start < nblk; start += blockDim.x) {
    int64_t i = start + tid;
    int64_t v = (i < nblk) ? blocksum[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    int64_t carry = carry_s;
    if (i < nblk) blocksum[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) blocksum[nblk] = carry_s;
}

__global__ void scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ blocksum,
                                int64_t nblk) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += blocksum[i / SCAN_TILE];
  if (i == 0) out[n] = blocksum[nblk];
}

// ------------------------------------------------------------------------ symmetrise
enum { SYM_PLUS = 0, SYM_MULT = 1, SYM_MNN = 2, SYM_NONE = 3 };

__device__ __forceinline__ double sym_combine(int mode, double theta, double w, double wr) {
  if (mode == SYM_PLUS) return (w + wr) / 2;
  if (mode == SYM_MULT) return w * wr;
  double lo = fmin(w, wr), hi = fmax(w, wr);
  return theta * lo + (1 - theta) * hi;
}

// position of column `c` in the column-sorted slice idx[lo, hi), or -1
__device__ __forceinline__ int64_t find_col(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t c) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t v = idx[mid];
    if (v < c) lo = mid + 1;
    else if (v > c) hi = mid;
    else return mid;
  }
  return -1;
}

constexpr int SYM_GROUP = 8;  // lanes cooperating on one row (raw rows hold ~10 edges)

template <bool FILL>
__global__ void __launch_bounds__(256) sym_pass_kernel(const int64_t* __restrict__ indptr,
                                                       const int32_t* __restrict__ idx,
                                                       const double* __restrict__ val, int64_t n, int mode,
                                                       double theta, int32_t* __restrict__ newlen,
                                                       int32_t* __restrict__ flags,
                                                       const int64_t* __restrict__ outptr,
                                                       int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_idx,
                                                       double* __restrict__ tmp_val) {
  const int sub = threadIdx.x % SYM_GROUP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SYM_GROUP;
  const bool active = row < n;
  int self_cnt = 0;
  bool asym = false;
  if (active) {
    const int64_t e0 = indptr[row], e1 = indptr[row + 1];
    for (int64_t e = e0 + sub; e < e1; e += SYM_GROUP) {
      const int32_t j = idx[e];
      const double w = val[e];
      double wr = w;
      int64_t pos = e;
      if (j != row) {
        pos = find_col(idx, indptr[j], indptr[j + 1], (int32_t)row);
        wr = pos >= 0 ? val[pos] : 0.0;
      }
      if (mode == SYM_NONE) {
        if (w - wr > 1e-5) asym = true;
        continue;
      }
      const double s = sym_combine(mode, theta, w, wr);
      if (s != 0.0) {
        if (FILL) {
          int64_t o = outptr[row] + atomicAdd(cursor + row, 1);
          tmp_idx[o] = j; tmp_val[o] = s;
          if (pos < 0) {
            int64_t o2 = outptr[j] + atomicAdd(cursor + j, 1);
            tmp_idx[o2] = (int32_t)row; tmp_val[o2] = s;
          }
        } else {
          ++self_cnt;
          if (pos < 0) atomicAdd(newlen + j, 1);
        }
      }
    }
  }
  if (!FILL) {
#pragma unroll
    for (int off = SYM_GROUP / 2; off > 0; off >>= 1) self_cnt += __shfl_xor_sync(0xffffffffu, self_cnt, off);
    if (active && sub == 0 && self_cnt) atomicAdd(newlen + row, self_cnt);
    if (asym) atomicOr(flags, 1);
  }
}

// ------------------------------------------------- per-row sort + normalise (warp per row)
constexpr int FIN_WARPS = 4, FIN_CAP = 64;

// SORT: rows of (tmp_idx,tmp_val) are unsorted -> sort by column.  Writes K (idx,val), P = val/rowsum,
// degree = rowsum; flags bit 1 set when a row of a square matrix has no diagonal entry.
template <bool SORT>
__global__ void __launch_bounds__(FIN_WARPS * 32) row_finalize_kernel(
    const int64_t* __restrict__ ptr, const int32_t* __restrict__ tmp_idx, const double* __restrict__ tmp_val,
    int64_t n, int32_t* __restrict__ out_idx, double* __restrict__ out_val, double* __restrict__ p_val,
    double* __restrict__ degree, int32_t* __restrict__ flags, int check_diag) {
  __shared__ int32_t ks[FIN_WARPS][FIN_CAP];
  __shared__ double vs[FIN_WARPS][FIN_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * FIN_WARPS + warp;
  if (row >= n) return;
  const int64_t p0 = ptr[row], p1 = ptr[row + 1];
  const int64_t L = p1 - p0;
  auto sync = [] { __syncwarp(); };
  double sum = 0.0;
  bool has_diag = false;
  if (!SORT) {
    for (int64_t t = lane; t < L; t += 32) {
      sum += fabs(tmp_val[p0 + t]);
      has_diag |= (tmp_idx[p0 + t] == row);
    }
  } else if (L <= 32) {
    // the common case (a kNN kernel row): one entry per lane, sorted by column in registers
    int32_t c = 0x7fffffff;
    double w = 0.0;
    if (lane < L) { c = tmp_idx[p0 + lane]; w = tmp_val[p0 + lane]; }
    warp_sort32<int32_t, double>(c, w, lane);
    if (lane < L) {
      out_idx[p0 + lane] = c;
      out_val[p0 + lane] = w;
      sum += fabs(w);
      has_diag |= (c == row);
    }
  } else if (L <= FIN_CAP) {
    int32_t* k = ks[warp];
    double* v = vs[warp];
    int np2 = 2;
    while (np2 < L) np2 <<= 1;
    for (int t = lane; t < np2; t += 32) {
      if (t < L) { k[t] = tmp_idx[p0 + t]; v[t] = tmp_val[p0 + t]; }
      else { k[t] = 0x7fffffff; v[t] = 0.0; }
    }
    __syncwarp();
    GTB_BITONIC_SORT(k, v, np2, lane, 32, sync, int32_t, double);
    for (int t = lane; t < L; t += 32) {
      out_idx[p0 + t] = k[t];
      out_val[p0 + t] = v[t];
      sum += fabs(v[t]);
      has_diag |= (k[t] == row);
    }
  } else {
    // long (hub) rows: rank by counting, columns are unique within a row
    for (int64_t t = lane; t < L; t += 32) {
      const int32_t c = tmp_idx[p0 + t];
      const double w = tmp_val[p0 + t];
      int64_t rank = 0;
      for (int64_t u = 0; u < L; ++u) rank += (tmp_idx[p0 + u] < c);
      out_idx[p0 + rank] = c;
      out_val[p0 + rank] = w;
      sum += fabs(w);
      has_diag |= (c == row);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  has_diag = __any_sync(0xffffffffu, has_diag);
  __syncwarp();
  if (p_val) {
    const double* src = SORT ? out_val : tmp_val;
    if (SORT && L > FIN_CAP) __threadfence_block();
    for (int64_t t = lane; t < L; t += 32) {
      double w = src[p0 + t];
      p_val[p0 + t] = (sum != 0.0) ? w / sum : w;
    }
  }
  if (lane == 0) {
    if (degree) degree[row] = sum;
    if (check_diag && !has_diag) atomicOr(flags, 2);
  }
}

__global__ void anisotropy_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                  double* __restrict__ val, const double* __restrict__ deg, double alpha, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const double di = deg[row];
  for (int64_t e = indptr[row] + lane; e < indptr[row + 1]; e += 32)
    val[e] = val[e] / pow(di * deg[idx[e]], alpha);
}

// ------------------------------------------------------------------ MNN block assembly
__global__ void block_count_kernel(const int64_t* __restrict__ indptr, int64_t nb, const int32_t* __restrict__ row_map,
                                   int32_t* __restrict__ rowlen) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) rowlen[row_map[i]] += (int32_t)(indptr[i + 1] - indptr[i]);
}

__global__ void __launch_bounds__(256) block_fill_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t nb,
    const int32_t* __restrict__ row_map, const int32_t* __restrict__ col_map, const double* __restrict__ within,
    const double* __restrict__ between, double beta, const int64_t* __restrict__ outptr,
    int32_t* __restrict__ cursor, int32_t* __restrict__ out_idx, double* __restrict__ out_val) {
  const int sub = threadIdx.x % SYM_GROUP;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SYM_GROUP;
  if (i >= nb) return;  // whole 8-lane group leaves together
  const int64_t e0 = indptr[i], e1 = indptr[i + 1];
  const int32_t grow = row_map[i];
  // every lane reads the cursor before the group leader advances it
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) / SYM_GROUP * SYM_GROUP);
  int32_t cur = cursor[grow];
  __syncwarp(gmask);
  double scale = 1.0;
  if (within) scale = fmin(1.0, within[i] / between[i]) * beta;  // graphs.py:1921-1925
  const int64_t o0 = outptr[grow] + cur;
  for (int64_t e = e0 + sub; e < e1; e += SYM_GROUP) {
    out_idx[o0 + (e - e0)] = col_map[idx[e]];
    out_val[o0 + (e - e0)] = val[e] * scale;
  }
  if (sub == 0) cursor[grow] = cur + (int32_t)(e1 - e0);
}

// ------------------------------------------- row-shard merge (multi-GPU symmetrisation)
// Row r of A = this rank's raw kernel rows, row r of B = the transposed edges routed to this rank by
// the all-to-all (both column-sorted).  Two-pointer merge per row: s(w, w') with w' = 0 where absent.
// FILL = false: counts the non-zero results; FILL = true: writes K (sorted), P = K / rowsum and degree.
template <bool FILL>
__global__ void sym_merge_rows_kernel(const int64_t* __restrict__ pa, const int32_t* __restrict__ ia,
                                      const double* __restrict__ va, const int64_t* __restrict__ pb,
                                      const int32_t* __restrict__ ib, const double* __restrict__ vb, int64_t n_rows,
                                      int mode, double theta, int32_t* __restrict__ newlen,
                                      const int64_t* __restrict__ outptr, int32_t* __restrict__ out_idx,
                                      double* __restrict__ out_val, double* __restrict__ p_val,
                                      double* __restrict__ degree) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  int64_t a = pa[row], a1 = pa[row + 1], b = pb[row], b1 = pb[row + 1];
  int64_t o = FILL ? outptr[row] : 0;
  const int64_t o0 = o;
  int cnt = 0;
  double sum = 0.0;
  while (a < a1 || b < b1) {
    const int32_t ca = (a < a1) ? ia[a] : 0x7fffffff, cb = (b < b1) ? ib[b] : 0x7fffffff;
    const int32_t c = ca < cb ? ca : cb;
    const double w = (ca == c) ? va[a] : 0.0, wr = (cb == c) ? vb[b] : 0.0;
    a += (ca == c);
    b += (cb == c);
    const double sv = sym_combine(mode, theta, w, wr);
    if (sv != 0.0) {
      if (FILL) { out_idx[o] = c; out_val[o] = sv; ++o; sum += fabs(sv); }
      else ++cnt;
    }
  }
  if (!FILL) { newlen[row] = cnt; return; }
  if (p_val) for (int64_t e = o0; e < o; ++e) p_val[e] = (sum != 0.0) ? out_val[e] / sum : out_val[e];
  if (degree) degree[row] = sum;
}

// dense[row][idx[e]] = val[e]; the output was zero-filled by the caller (cudaMemsetAsync)
__global__ void csr_to_dense_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                    const double* __restrict__ val, int64_t n_rows, int64_t n_cols,
                                    double* __restrict__ out) {
  const int sub = threadIdx.x % SYM_GROUP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SYM_GROUP;
  if (row >= n_rows) return;
  for (int64_t e = indptr[row] + sub; e < indptr[row + 1]; e += SYM_GROUP) out[row * n_cols + idx[e]] = val[e];
}

__global__ void cast_indptr_kernel(const int64_t* __restrict__ in, int64_t n1, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) out[i] = (int32_t)in[i];
}

}  // namespace

extern "C" int64_t gtb_scan_ws_elems(int64_t n) { return gtb_cdiv(n > 0 ? n : 1, SCAN_TILE) + 2; }

extern "C" int gtb_exclusive_scan(const int32_t* in, int64_t n, int64_t* out, int64_t* ws, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty scan");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t nblk = gtb_cdiv(n, SCAN_TILE);
  scan_block_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(in, n, out, ws);
  GTB_CHECK_LAUNCH();
  scan_top_kernel<<<1, 1024, 0, st>>>(ws, nblk);
  GTB_CHECK_LAUNCH();
  scan_add_kernel<<<(unsigned)gtb_cdiv(n, 256), 256, 0, st>>>(out, n, ws, nblk);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_cast_indptr(const int64_t* in, int64_t n1, int32_t* out, void* stream) {
  cast_indptr_kernel<<<(unsigned)gtb_cdiv(n1, 256), 256, 0, (cudaStream_t)stream>>>(in, n1, out);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_sym_count(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n, int mode,
                             double theta, int32_t* newlen, int32_t* flags, void* stream) {
  GTB_CHECK_ARG(n > 0 && mode >= 0 && mode <= 3, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(newlen, 0, sizeof(int32_t) * n, st));
  sym_pass_kernel<false><<<(unsigned)gtb_cdiv(n * SYM_GROUP, 256), 256, 0, st>>>(
      indptr, idx, val, n, mode, theta, newlen, flags, nullptr, nullptr, nullptr, nullptr);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_sym_fill(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n, int mode,
                            double theta, const int64_t* outptr, int32_t* cursor, int32_t* tmp_idx,
                            double* tmp_val, void* stream) {
  GTB_CHECK_ARG(n > 0 && mode >= 0 && mode <= 2, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, st));
  sym_pass_kernel<true><<<(unsigned)gtb_cdiv(n * SYM_GROUP, 256), 256, 0, st>>>(
      indptr, idx, val, n, mode, theta, nullptr, nullptr, outptr, cursor, tmp_idx, tmp_val);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_row_finalize(const int64_t* ptr, const int32_t* tmp_idx, const double* tmp_val, int64_t n,
                                int sort, int32_t* out_idx, double* out_val, double* p_val, double* degree,
                                int32_t* flags, int check_diag, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)gtb_cdiv(n, FIN_WARPS);
  if (sort)
    row_finalize_kernel<true><<<grid, FIN_WARPS * 32, 0, st>>>(ptr, tmp_idx, tmp_val, n, out_idx, out_val, p_val,
                                                               degree, flags, check_diag);
  else
    row_finalize_kernel<false><<<grid, FIN_WARPS * 32, 0, st>>>(ptr, tmp_idx, tmp_val, n, out_idx, out_val, p_val,
                                                                degree, flags, check_diag);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_block_count(const int64_t* indptr, int64_t nb, const int32_t* row_map, int32_t* rowlen,
                               void* stream) {
  GTB_CHECK_ARG(nb > 0, "empty block");
  block_count_kernel<<<(unsigned)gtb_cdiv(nb, 256), 256, 0, (cudaStream_t)stream>>>(indptr, nb, row_map, rowlen);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_block_fill(const int64_t* indptr, const int32_t* idx, const double* val, int64_t nb,
                              const int32_t* row_map, const int32_t* col_map, const double* within,
                              const double* between, double beta, const int64_t* outptr, int32_t* cursor,
                              int32_t* out_idx, double* out_val, void* stream) {
  GTB_CHECK_ARG(nb > 0, "empty block");
  block_fill_kernel<<<(unsigned)gtb_cdiv(nb * SYM_GROUP, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, idx, val, nb, row_map, col_map, within, between, beta, outptr, cursor, out_idx, out_val);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_sym_merge_count(const int64_t* pa, const int32_t* ia, const double* va, const int64_t* pb,
                                   const int32_t* ib, const double* vb, int64_t n_rows, int mode, double theta,
                                   int32_t* newlen, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && mode >= 0 && mode <= 2, "bad arguments");
  sym_merge_rows_kernel<false><<<(unsigned)gtb_cdiv(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
      pa, ia, va, pb, ib, vb, n_rows, mode, theta, newlen, nullptr, nullptr, nullptr, nullptr, nullptr);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_sym_merge_fill(const int64_t* pa, const int32_t* ia, const double* va, const int64_t* pb,
                                  const int32_t* ib, const double* vb, int64_t n_rows, int mode, double theta,
                                  const int64_t* outptr, int32_t* out_idx, double* out_val, double* p_val,
                                  double* degree, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && mode >= 0 && mode <= 2, "bad arguments");
  sym_merge_rows_kernel<true><<<(unsigned)gtb_cdiv(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(
      pa, ia, va, pb, ib, vb, n_rows, mode, theta, nullptr, outptr, out_idx, out_val, p_val, degree);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_csr_to_dense(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n_rows,
                                int64_t n_cols, double* out, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && n_cols > 0, "empty matrix");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)n_rows * (size_t)n_cols, st));
  csr_to_dense_kernel<<<(unsigned)gtb_cdiv(n_rows * SYM_GROUP, 256), 256, 0, st>>>(indptr, idx, val, n_rows, n_cols, out);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_anisotropy(const int64_t* indptr, const int32_t* idx, double* val, const double* deg,
                              double alpha, int64_t n, void* stream) {
  anisotropy_kernel<<<(unsigned)gtb_cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(indptr, idx, val, deg, alpha, n);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
