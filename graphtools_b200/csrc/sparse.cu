// Sparse helpers shared by every stage that sizes or finishes a CSR: the single-pass exclusive scan, per-row
// normalisation (P = K / rowsum, degree, diagonal check), anisotropy, MNN block assembly, densification.
// The symmetrisation itself (sort-based transpose + merge) lives in symm.cu.
//
// Replaces sklearn normalize(K, "l1") (reference graphtools/base.py:645), kernel_degree (base.py:648-666),
// apply_anisotropy (base.py:579-592), the np.cumsum of _build_csr_from_neighbors (graphs.py:519-551) and
// matrix.set_submatrix (matrix.py:49-51).
#include "common.cuh"
#include "gtb200.h"

namespace {

// ------------------------------------------------------------------------------ scan
// Single-pass exclusive scan (int32 -> int64) with decoupled look-back: every tile publishes its aggregate, then
// its inclusive prefix, in one 64-bit word (2 flag bits : 62 value bits) and resolves its own prefix by walking
// back over its predecessors' words.  Tile ids are handed out by an atomic ticket, so a tile's predecessors have
// always started -- no dependence on block scheduling order.  ws[0] = ticket, ws[1 + t] = state of tile t.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_AGG = 1ull << 62, SCAN_PRE = 2ull << 62, SCAN_VAL = (1ull << 62) - 1;

__global__ void __launch_bounds__(SCAN_THREADS) scan_lookback_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                     int64_t* __restrict__ out,
                                                                     unsigned long long* __restrict__ ws) {
  __shared__ int64_t warp_tot[SCAN_THREADS / 32];
  __shared__ int64_t tile_s, prefix_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) tile_s = (int64_t)atomicAdd(ws, 1ull);
  __syncthreads();
  const int64_t tile = tile_s;
  unsigned long long* state = ws + 1;
  const int64_t base = tile * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
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
    const int64_t o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int64_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const int64_t x = warp_tot[w];
    if (w < warp) wbase += x;
    total += x;
  }
  if (warp == 0) {
    int64_t excl = 0;
    if (tile > 0) {
      if (lane == 0) atomicExch(state + tile, SCAN_AGG | (unsigned long long)total);
      int64_t look = tile - 1;
      while (true) {
        const int64_t t = look - lane;
        unsigned long long word = SCAN_PRE;                       // tiles before 0: prefix 0
        if (t >= 0) {
          do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(word) : "l"(state + t) : "memory");
          } while ((word >> 62) == 0);
        }
        const unsigned pre = __ballot_sync(0xffffffffu, (word >> 62) == 2);
        const int first = pre ? (__ffs(pre) - 1) : 32;             // nearest predecessor that already has a prefix
        int64_t part = (lane <= first) ? (int64_t)(word & SCAN_VAL) : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        excl += part;
        if (pre) break;
        look -= 32;
      }
    }
    if (lane == 0) {
      atomicExch(state + tile, SCAN_PRE | (unsigned long long)(excl + total));
      prefix_s = excl;
    }
  }
  __syncthreads();
  int64_t run = prefix_s + wbase + incl - tsum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;   // the thread that owns the last element
}

// ------------------------------------------------------------------------ asymmetry check
// kernel_symm=None: the reference warns when max(K - K^T) > 1e-5 (base.py:551-552).  Every edge binary-searches
// row j for column i (rows are column-sorted); an absent reverse edge counts as 0.
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

__global__ void __launch_bounds__(256) asym_check_kernel(const int64_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ idx,
                                                         const double* __restrict__ val, int64_t n,
                                                         int32_t* __restrict__ flags) {
  const int sub = threadIdx.x % SYM_GROUP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / SYM_GROUP;
  if (row >= n) return;
  bool asym = false;
  for (int64_t e = indptr[row] + sub; e < indptr[row + 1]; e += SYM_GROUP) {
    const int32_t j = idx[e];
    if (j == row) continue;
    const int64_t pos = find_col(idx, indptr[j], indptr[j + 1], (int32_t)row);
    const double wr = pos >= 0 ? val[pos] : 0.0;
    if (val[e] - wr > 1e-5) asym = true;
  }
  if (asym) atomicOr(flags, 1);
}

// ------------------------------------------------- per-row normalise (warp per row)
constexpr int FIN_WARPS = 4;

// P = val / rowsum (sklearn normalize(K, "l1"): rows summing to zero are left as they are), degree = rowsum;
// flags bit 1 set when a row of a square matrix has no diagonal entry.
__global__ void __launch_bounds__(FIN_WARPS * 32) row_finalize_kernel(
    const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t n,
    double* __restrict__ p_val, double* __restrict__ degree, int32_t* __restrict__ flags, int check_diag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * FIN_WARPS + warp;
  if (row >= n) return;
  const int64_t p0 = ptr[row], p1 = ptr[row + 1];
  const int64_t L = p1 - p0;
  double sum = 0.0;
  bool has_diag = false;
  for (int64_t t = lane; t < L; t += 32) {
    sum += fabs(val[p0 + t]);
    has_diag |= (idx[p0 + t] == row);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  has_diag = __any_sync(0xffffffffu, has_diag);
  if (p_val) {
    for (int64_t t = lane; t < L; t += 32) {
      const double w = val[p0 + t];
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
  const int64_t nblk = gtb_cdiv(n, SCAN_TILE);
  GTB_CUDA(cudaMemsetAsync(ws, 0, sizeof(int64_t) * (size_t)(nblk + 1), st));
  scan_lookback_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(in, n, out, reinterpret_cast<unsigned long long*>(ws));
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_cast_indptr(const int64_t* in, int64_t n1, int32_t* out, void* stream) {
  cast_indptr_kernel<<<(unsigned)gtb_cdiv(n1, 256), 256, 0, (cudaStream_t)stream>>>(in, n1, out);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_asym_check(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n,
                              int32_t* flags, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  asym_check_kernel<<<(unsigned)gtb_cdiv(n * SYM_GROUP, 256), 256, 0, (cudaStream_t)stream>>>(indptr, idx, val, n,
                                                                                            flags);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_row_finalize(const int64_t* ptr, const int32_t* idx, const double* val, int64_t n, double* p_val,
                                double* degree, int32_t* flags, int check_diag, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  row_finalize_kernel<<<(unsigned)gtb_cdiv(n, FIN_WARPS), FIN_WARPS * 32, 0, (cudaStream_t)stream>>>(
      ptr, idx, val, n, p_val, degree, flags, check_diag);
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
