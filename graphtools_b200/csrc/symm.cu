// K4: sort-based sparse transpose + symmetrisation (+, *, mnn) + row normalisation into the diffusion
// operator, in streaming passes over the CSR -- and the bucketing kernels of the multi-GPU edge exchange.
//
// Replaces scipy's csr binops behind BaseGraph.symmetrize_kernel (reference graphtools/base.py:557-577,
// matrix.py:16-29), sklearn normalize(K, "l1") (base.py:645), kernel_degree (base.py:648-666) and the diagonal
// check of BaseGraph._build_kernel (base.py:553-554).
//
//   transpose_count   histogram of the column indices                          (thread per edge, one atomic each)
//   scan              row pointers of R^T (sparse.cu, single pass), cast to 32-bit cursors
//   transpose_scatter every edge (i, j, w) -> one 16-byte record {i, j, w} in row j of R^T: ONE atomic on the row's
//                     cursor (which already holds the row's start) + ONE 16-byte store per edge
//   rec_sort_rows     rows that are too long for the register path of the merge are ordered by column, in place
//   sym_merge<count>  |row i of R  UNION  row i of R^T| under the merge rule     (warp per row)
//   scan
//   sym_merge<fill>   K row (column-sorted), P = K / rowsum, degree, diagonal flag, one sweep
//
// The merge takes A = a row of R (column-sorted CSR) and T = the matching row of R^T as records in ARRIVAL order.
// Rows with |A| + |T| <= 32 (every row of a kNN kernel on low-intrinsic-dimension data) live one element per lane:
// mutual edges are found with one match.any on the column, output positions by counting -- no sort at all, and the
// result does not depend on the order in which the atomics of the scatter landed (columns are unique within A and
// within T).  Longer rows are sorted first and merged with two pointers.  K, P and the degree vector are therefore
// bit-reproducible and identical between the single-GPU build and the row-sharded multi-GPU build (same kernels,
// same per-row order).  All three merge rules are symmetric functions s(w, w'), hence K is bitwise symmetric.
#include "common.cuh"
#include "gtb200.h"

namespace {

enum { SYM_PLUS = 0, SYM_MULT = 1, SYM_MNN = 2 };

__device__ __forceinline__ double sym_combine(int mode, double theta, double w, double wr) {
  if (mode == SYM_PLUS) return (w + wr) / 2;
  if (mode == SYM_MULT) return w * wr;
  // scipy: theta * K.minimum(K.T) + (1 - theta) * K.maximum(K.T) -- two rounded products, one rounded sum
  const double lo = fmin(w, wr), hi = fmax(w, wr);
  return __dadd_rn(__dmul_rn(theta, lo), __dmul_rn(1 - theta, hi));
}

constexpr int GRP = 8;  // lanes cooperating on one raw row (~9 edges)

// ------------------------------------------------------------------------------ transpose
struct __align__(16) EdgeRec { int32_t i, j; double w; };     // edge (i, j, w): entry (j, i) of the transposed matrix

__device__ __forceinline__ void store_rec(EdgeRec* dst, int32_t i, int32_t j, double w) {
  int4 v;
  v.x = i; v.y = j;
  v.z = __double2loint(w); v.w = __double2hiint(w);
  *reinterpret_cast<int4*>(dst) = v;
}

__global__ void __launch_bounds__(256) transpose_count_kernel(const int32_t* __restrict__ idx, int64_t nnz,
                                                              int32_t col0, int32_t* __restrict__ cnt) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(cnt + (idx[e] - col0), 1);
}

// cursor[j] = start of row j of the transposed matrix on entry (32-bit copy of the scanned row pointers); every edge
// takes the next slot of its row with one atomic and writes its record with one 16-byte store
__global__ void __launch_bounds__(256) transpose_scatter_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
    int64_t n_rows, int32_t row0, int32_t col0, int32_t* __restrict__ cursor, EdgeRec* __restrict__ t_rec) {
  const int sub = threadIdx.x % GRP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GRP;
  if (row >= n_rows) return;
  const int64_t e1 = indptr[row + 1];
  for (int64_t e = indptr[row] + sub; e < e1; e += GRP) {
    const int32_t j = idx[e];
    const int32_t o = atomicAdd(cursor + (j - col0), 1);
    store_rec(t_rec + o, (int32_t)row + row0, j, val[e]);
  }
}

// the same two steps for a list of packed edge records {i, j, w} (what the all-to-all delivers)
__global__ void __launch_bounds__(256) records_count_kernel(const EdgeRec* __restrict__ rec, int64_t k, int32_t col0,
                                                            int32_t* __restrict__ cnt) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < k; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(cnt + (rec[e].j - col0), 1);
}

__global__ void __launch_bounds__(256) records_scatter_kernel(const EdgeRec* __restrict__ rec, int64_t k, int32_t col0,
                                                              int32_t* __restrict__ cursor,
                                                              EdgeRec* __restrict__ t_rec) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < k; e += (int64_t)gridDim.x * blockDim.x) {
    const int4 r = reinterpret_cast<const int4*>(rec)[e];
    const int32_t o = atomicAdd(cursor + (r.y - col0), 1);
    reinterpret_cast<int4*>(t_rec)[o] = r;
  }
}

// ------------------------------------------------------------------------------ segmented sort (in place)
// Rows of a CSR ordered by column; columns are unique within a row.  Storage policies: separate index / value arrays
// (MNN block assembly) or 16-byte edge records keyed by .i (rows of a transposed matrix).  Tiers: <= 32 entries: one
// per lane, rank by counting (3 instructions per entry, no network); <= SORT_WARP_CAP: warp bitonic in shared
// memory; longer rows are left to csr_sort_long_kernel (block per row: shared memory up to SORT_BLOCK_CAP, global
// memory beyond).  `pa` (optional): rows whose length plus the length of the matching row of `pa` stays <= min_total
// are skipped -- the merge handles those unsorted.
constexpr int SORT_WARPS = 4, SORT_WARP_CAP = 256, SORT_BLOCK_CAP = 4096, SORT_BLOCK_THREADS = 256;

struct SoAStore {
  int32_t* idx; double* val;
  __device__ __forceinline__ int32_t key(int64_t p) const { return idx[p]; }
  __device__ __forceinline__ double pay(int64_t p) const { return val[p]; }
  __device__ __forceinline__ int32_t aux(int64_t) const { return 0; }
  __device__ __forceinline__ void put(int64_t p, int32_t k, double v, int32_t) const { idx[p] = k; val[p] = v; }
};
struct RecStore {
  EdgeRec* rec;
  __device__ __forceinline__ int32_t key(int64_t p) const { return rec[p].i; }
  __device__ __forceinline__ double pay(int64_t p) const { return rec[p].w; }
  __device__ __forceinline__ int32_t aux(int64_t p0) const { return rec[p0].j; }      // constant within a row
  __device__ __forceinline__ void put(int64_t p, int32_t k, double v, int32_t j) const { store_rec(rec + p, k, j, v); }
};

__device__ __forceinline__ bool sort_skips(const int64_t* __restrict__ pa, int64_t row, int64_t L, int min_total) {
  const int64_t la = pa ? (pa[row + 1] - pa[row]) : 0;
  return L <= 1 || la + L <= min_total;
}

template <typename Store>
__global__ void __launch_bounds__(SORT_WARPS * 32) csr_sort_rows_kernel(const int64_t* __restrict__ ptr, Store st,
                                                                        int64_t n, const int64_t* __restrict__ pa,
                                                                        int min_total, int32_t* __restrict__ has_long) {
  __shared__ int32_t ks[SORT_WARPS][SORT_WARP_CAP];
  __shared__ double vs[SORT_WARPS][SORT_WARP_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * SORT_WARPS + warp;
  if (row >= n) return;
  const int64_t p0 = ptr[row];
  const int64_t L = ptr[row + 1] - p0;
  if (sort_skips(pa, row, L, min_total)) return;
  if (L <= 32) {
    int32_t c = 0x7fffffff;
    double w = 0.0;
    const int32_t aux = st.aux(p0);
    if (lane < L) { c = st.key(p0 + lane); w = st.pay(p0 + lane); }
    int rank = 0;
    for (int k = 0; k < (int)L; ++k) rank += (__shfl_sync(0xffffffffu, c, k) < c);
    __syncwarp();
    if (lane < L) st.put(p0 + rank, c, w, aux);
  } else if (L <= SORT_WARP_CAP) {
    int32_t* k = ks[warp];
    double* v = vs[warp];
    const int32_t aux = st.aux(p0);
    int np2 = 64;
    while (np2 < L) np2 <<= 1;
    for (int t = lane; t < np2; t += 32) {
      if (t < L) { k[t] = st.key(p0 + t); v[t] = st.pay(p0 + t); }
      else { k[t] = 0x7fffffff; v[t] = 0.0; }
    }
    __syncwarp();
    auto sync = [] { __syncwarp(); };
    GTB_BITONIC_SORT(k, v, np2, lane, 32, sync, int32_t, double);
    for (int t = lane; t < L; t += 32) st.put(p0 + t, k[t], v[t], aux);
  } else if (lane == 0) {
    *has_long = 1;
  }
}

template <typename Store>
__global__ void __launch_bounds__(SORT_BLOCK_THREADS) csr_sort_long_kernel(const int64_t* __restrict__ ptr, Store st,
                                                                           int64_t n,
                                                                           const int32_t* __restrict__ has_long) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  double* v = reinterpret_cast<double*>(sort_smem);                    // [SORT_BLOCK_CAP]
  int32_t* k = reinterpret_cast<int32_t*>(v + SORT_BLOCK_CAP);         // [SORT_BLOCK_CAP]
  __shared__ unsigned int long_mask[SORT_BLOCK_THREADS / 32];
  if (*has_long == 0) return;
  const int tid = threadIdx.x;
  auto sync = [] { __syncthreads(); };
  for (int64_t base = (int64_t)blockIdx.x * SORT_BLOCK_THREADS; base < n; base += (int64_t)gridDim.x * SORT_BLOCK_THREADS) {
    const int64_t r = base + tid;
    const bool is_long = (r < n) && (ptr[r + 1] - ptr[r] > SORT_WARP_CAP);
    const unsigned m = __ballot_sync(0xffffffffu, is_long);
    if ((tid & 31) == 0) long_mask[tid >> 5] = m;
    __syncthreads();
    for (int w = 0; w < SORT_BLOCK_THREADS / 32; ++w) {
      unsigned mm = long_mask[w];
      while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        const int64_t row = base + w * 32 + b;
        const int64_t p0 = ptr[row];
        const int64_t L = ptr[row + 1] - p0;
        const int32_t aux = st.aux(p0);
        __syncthreads();
        if (L <= SORT_BLOCK_CAP) {
          int np2 = 512;
          while (np2 < L) np2 <<= 1;
          for (int t = tid; t < np2; t += SORT_BLOCK_THREADS) {
            if (t < L) { k[t] = st.key(p0 + t); v[t] = st.pay(p0 + t); }
            else { k[t] = 0x7fffffff; v[t] = 0.0; }
          }
          __syncthreads();
          GTB_BITONIC_SORT(k, v, np2, tid, SORT_BLOCK_THREADS, sync, int32_t, double);
          for (int t = tid; t < L; t += SORT_BLOCK_THREADS) st.put(p0 + t, k[t], v[t], aux);
          __syncthreads();
        } else {
          // hub rows beyond the shared-memory tile: bitonic network straight on global memory, in the form whose
          // compare-exchanges all point the same way (first step of a stage pairs t with its mirror t ^ (kk - 1)):
          // virtual +infinity padding past L then never has to move, so no scratch row is needed
          int64_t np2 = SORT_BLOCK_CAP;
          while (np2 < L) np2 <<= 1;
          for (int64_t kk = 2; kk <= np2; kk <<= 1) {
            for (int64_t j = kk >> 1; j > 0; j >>= 1) {
              for (int64_t t = tid; t < L; t += SORT_BLOCK_THREADS) {
                const int64_t q = (j == (kk >> 1)) ? (t ^ (kk - 1)) : (t ^ j);
                if (q > t && q < L) {
                  const int32_t a = st.key(p0 + t), b2 = st.key(p0 + q);
                  if (a > b2) {
                    const double va = st.pay(p0 + t), vb = st.pay(p0 + q);
                    st.put(p0 + t, b2, vb, aux);
                    st.put(p0 + q, a, va, aux);
                  }
                }
              }
              __threadfence_block();
              __syncthreads();
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

template <typename Store>
int launch_sort(const int64_t* ptr, Store st, int64_t n, const int64_t* pa, int min_total, int32_t* has_long,
                cudaStream_t stream) {
  GTB_CUDA(cudaMemsetAsync(has_long, 0, sizeof(int32_t), stream));
  csr_sort_rows_kernel<Store><<<(unsigned)gtb_cdiv(n, SORT_WARPS), SORT_WARPS * 32, 0, stream>>>(ptr, st, n, pa,
                                                                                               min_total, has_long);
  GTB_CHECK_LAUNCH();
  const size_t smem = (size_t)SORT_BLOCK_CAP * 12;
  GTB_CUDA(cudaFuncSetAttribute(csr_sort_long_kernel<Store>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = gtb_cdiv(n, SORT_BLOCK_THREADS);
  csr_sort_long_kernel<Store><<<(unsigned)(blocks < 148 * 4 ? blocks : 148 * 4), SORT_BLOCK_THREADS, smem, stream>>>(
      ptr, st, n, has_long);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

// ------------------------------------------------------------------------------ merge
// Row r of A = raw kernel rows [pa, ia, va] (column-sorted, columns global), row r of T = records [pt, tr] of the
// transposed matrix (entry column = .i, value = .w): in arrival order when |A| + |T| <= 32, column-sorted otherwise
// (rec_sort_rows).  FILL = false: newlen[r] = number of non-zero results; FILL = true: K row, P = K / sum|K|, degree,
// flags bit 1 when the row has no diagonal entry (global row id = row0 + r).
constexpr int MRG_WARPS = 8;
constexpr int MRG_REG = 32;          // longest row handled one element per lane

template <bool FILL>
__global__ void __launch_bounds__(MRG_WARPS * 32) sym_merge_kernel(
    const int64_t* __restrict__ pa, const int32_t* __restrict__ ia, const double* __restrict__ va,
    const int64_t* __restrict__ pt, const EdgeRec* __restrict__ tr, int64_t n_rows, int32_t row0, int mode,
    double theta, int32_t* __restrict__ newlen, int32_t* __restrict__ worklist, int32_t* __restrict__ wl_count,
    const int64_t* __restrict__ outptr, int32_t* __restrict__ out_idx, double* __restrict__ out_val,
    double* __restrict__ p_val, double* __restrict__ degree, int32_t* __restrict__ flags) {
  __shared__ double sv[MRG_WARPS][32];
  __shared__ int32_t sc[MRG_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * MRG_WARPS + warp;
  if (row >= n_rows) return;
  const int64_t a0 = pa[row], t0 = pt[row];
  const int la = (int)(pa[row + 1] - a0), lt = (int)(pt[row + 1] - t0);
  const int L = la + lt;
  const int32_t grow = (int32_t)row + row0;
  if (L <= MRG_REG) {
    // ---- the common case (a kNN kernel row and its transpose): one element per lane, A in lanes [0, la), T in
    //      lanes [la, L) in arrival order.  Idle lanes carry distinct negative columns so they never match.
    const bool isA = lane < la, isT = (lane >= la) && (lane < L);
    int32_t c = -1 - lane;
    double w = 0.0;
    if (isA) { c = ia[a0 + lane]; w = va[a0 + lane]; }
    else if (isT) {
      const int4 r = reinterpret_cast<const int4*>(tr)[t0 + lane - la];
      c = r.x;
      w = __hiloint2double(r.w, r.z);
    }
    // columns are unique within A and within T: a column held by two lanes is a mutual edge.  (A broadcast loop --
    // match.any serialises on the number of distinct values and left the count pass latency-bound.)
    int partner = lane;
    for (int k = 0; k < L; ++k) {
      const int32_t ck = __shfl_sync(0xffffffffu, c, k);
      partner = (ck == c && k != lane) ? k : partner;
    }
    const bool mutual = partner != lane;
    const double w_other = __shfl_sync(0xffffffffu, w, partner);
    const double s = sym_combine(mode, theta, w, mutual ? w_other : 0.0);
    const bool emit = (isA || isT) && (s != 0.0) && !(mutual && isT);   // a mutual pair is emitted by its A copy
    const unsigned em = __ballot_sync(0xffffffffu, emit);
    const int cnt = __popc(em);
    if (!FILL) {
      if (lane == 0) newlen[row] = cnt;
      return;
    }
    // output position = number of emitted entries with a smaller column
    int rank = 0;
    for (int k = 0; k < L; ++k) {
      const int32_t ck = __shfl_sync(0xffffffffu, c, k);
      rank += (int)((em >> k) & 1u) & (int)(ck < c);
    }
    if (emit) { sv[warp][rank] = s; sc[warp][rank] = c; }
    __syncwarp();
    const bool on = lane < cnt;
    const double v = on ? sv[warp][lane] : 0.0;
    const int32_t cc = on ? sc[warp][lane] : -1;
    double sum = fabs(v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const bool has_diag = __any_sync(0xffffffffu, on && cc == grow);
    const int64_t o0 = outptr[row];
    if (on) {
      out_idx[o0 + lane] = cc;
      out_val[o0 + lane] = v;
      if (p_val) p_val[o0 + lane] = (sum != 0.0) ? v / sum : v;
    }
    if (lane == 0) {
      if (degree) degree[row] = sum;
      if (flags && !has_diag) atomicOr(flags, 2);
    }
    return;
  }
  // ---- rows with more than 32 entries (hub rows; every row of an isotropic, high-intrinsic-dimension data set)
  const int32_t* A = ia + a0;
  const EdgeRec* T = tr + t0;
  if (!FILL) {
    // count without sorted T: every T entry looks its column up in A (sorted).
    //   #emitted = sum_a [s(w_a, 0) != 0] + sum_t (found ? [s(w_a, w_t) != 0] - [s(w_a, 0) != 0] : [s(0, w_t) != 0])
    // The row is queued for the sort that the fill pass needs.
    int cnt = 0;
    for (int t = lane; t < la; t += 32) cnt += (sym_combine(mode, theta, va[a0 + t], 0.0) != 0.0);
    for (int t = lane; t < lt; t += 32) {
      const int32_t c = T[t].i;
      const double wt = T[t].w;
      int lo = 0, hi = la;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] < c) lo = mid + 1; else hi = mid;
      }
      if (lo < la && A[lo] == c) {
        const double wa = va[a0 + lo];
        cnt += (int)(sym_combine(mode, theta, wa, wt) != 0.0) - (int)(sym_combine(mode, theta, wa, 0.0) != 0.0);
      } else {
        cnt += (sym_combine(mode, theta, 0.0, wt) != 0.0);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (lane == 0) {
      newlen[row] = cnt;
      if (worklist) worklist[atomicAdd(wl_count, 1)] = (int32_t)row;
    }
    return;
  }
  // fill: T was sorted by rec_sort_worklist.  Sequential two-pointer merge by one lane (O(L), any length, fixed
  // order); the row sum runs in column order, which is sklearn's own summation order
  // (sparsefuncs_fast.pyx:_inplace_csr_row_normalize_l1)
  const int64_t o0 = outptr[row];
  int o = 0;
  double sum = 0.0;
  if (lane == 0) {
    bool has_diag = false;
    int a = 0, b = 0;
    while (a < la || b < lt) {
      const int32_t ca = (a < la) ? A[a] : 0x7fffffff, cb = (b < lt) ? T[b].i : 0x7fffffff;
      const int32_t c = ca < cb ? ca : cb;
      const double w = (ca == c) ? va[a0 + a] : 0.0, wr = (cb == c) ? T[b].w : 0.0;
      a += (ca == c);
      b += (cb == c);
      const double sv2 = sym_combine(mode, theta, w, wr);
      if (sv2 != 0.0) {
        out_idx[o0 + o] = c; out_val[o0 + o] = sv2; ++o;
        sum += fabs(sv2);
        has_diag |= (c == grow);
      }
    }
    if (degree) degree[row] = sum;
    if (flags && !has_diag) atomicOr(flags, 2);
  }
  o = __shfl_sync(0xffffffffu, o, 0);
  sum = __shfl_sync(0xffffffffu, sum, 0);
  __syncwarp();
  if (p_val)
    for (int e = lane; e < o; e += 32) { const double v = out_val[o0 + e]; p_val[o0 + e] = (sum != 0.0) ? v / sum : v; }
}

// Sort of the queued (long) record rows: one block per queue entry, shared-memory bitonic up to SORT_BLOCK_CAP entries,
// the uniform-direction global-memory network beyond.
__global__ void __launch_bounds__(SORT_BLOCK_THREADS) rec_sort_worklist_kernel(const int64_t* __restrict__ ptr,
                                                                               EdgeRec* __restrict__ rec,
                                                                               const int32_t* __restrict__ worklist,
                                                                               const int32_t* __restrict__ wl_count) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  double* v = reinterpret_cast<double*>(sort_smem);                    // [SORT_BLOCK_CAP]
  int32_t* k = reinterpret_cast<int32_t*>(v + SORT_BLOCK_CAP);         // [SORT_BLOCK_CAP]
  const int tid = threadIdx.x;
  const int nwork = *wl_count;
  auto sync = [] { __syncthreads(); };
  RecStore st{rec};
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int64_t row = worklist[wi];
    const int64_t p0 = ptr[row];
    const int64_t L = ptr[row + 1] - p0;
    if (L <= 1) continue;
    const int32_t aux = st.aux(p0);
    __syncthreads();
    if (L <= SORT_BLOCK_CAP) {
      int np2 = 2;
      while (np2 < L) np2 <<= 1;
      for (int t = tid; t < np2; t += SORT_BLOCK_THREADS) {
        if (t < L) { k[t] = st.key(p0 + t); v[t] = st.pay(p0 + t); }
        else { k[t] = 0x7fffffff; v[t] = 0.0; }
      }
      __syncthreads();
      GTB_BITONIC_SORT(k, v, np2, tid, SORT_BLOCK_THREADS, sync, int32_t, double);
      for (int t = tid; t < L; t += SORT_BLOCK_THREADS) st.put(p0 + t, k[t], v[t], aux);
    } else {
      int64_t np2 = SORT_BLOCK_CAP;
      while (np2 < L) np2 <<= 1;
      for (int64_t kk = 2; kk <= np2; kk <<= 1) {
        for (int64_t j = kk >> 1; j > 0; j >>= 1) {
          for (int64_t t = tid; t < L; t += SORT_BLOCK_THREADS) {
            const int64_t q = (j == (kk >> 1)) ? (t ^ (kk - 1)) : (t ^ j);
            if (q > t && q < L) {
              const int32_t a = st.key(p0 + t), b2 = st.key(p0 + q);
              if (a > b2) {
                const double va2 = st.pay(p0 + t), vb2 = st.pay(p0 + q);
                st.put(p0 + t, b2, vb2, aux);
                st.put(p0 + q, a, va2, aux);
              }
            }
          }
          __threadfence_block();
          __syncthreads();
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ edge routing (multi-GPU)
// Column owner under the block row partition: owner(j) = min(j / per, world - 1).  Rows are column-sorted, so the
// entries of a row bound for one owner are contiguous: cnt[o * m + r] = how many of row r go to rank o
// (destination-major, so one exclusive scan of cnt yields every send-buffer position).
__global__ void __launch_bounds__(256) route_count_kernel(const int64_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ idx, int64_t m, int32_t per,
                                                          int world, int32_t* __restrict__ cnt) {
  const int sub = threadIdx.x % GRP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GRP;
  if (row >= m) return;
  const int64_t e1 = indptr[row + 1];
  for (int64_t e = indptr[row] + sub; e < e1; e += GRP) {
    int o = idx[e] / per;
    o = o < world ? o : world - 1;
    atomicAdd(cnt + (int64_t)o * m + row, 1);
  }
}

__global__ void __launch_bounds__(256) route_fill_kernel(const int64_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ idx,
                                                         const double* __restrict__ val, int64_t m, int32_t row0,
                                                         int32_t per, int world, const int32_t* __restrict__ cnt,
                                                         const int64_t* __restrict__ pos, EdgeRec* __restrict__ send) {
  const int sub = threadIdx.x % GRP;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GRP;
  if (row >= m) return;
  const int64_t e0 = indptr[row], e1 = indptr[row + 1];
  for (int64_t e = e0 + sub; e < e1; e += GRP) {
    const int32_t j = idx[e];
    int o = j / per;
    o = o < world ? o : world - 1;
    int64_t before = 0;                           // entries of this row bound for lower ranks
    for (int q = 0; q < o; ++q) before += cnt[(int64_t)q * m + row];
    EdgeRec r;
    r.i = (int32_t)row + row0; r.j = j; r.w = val[e];
    send[pos[(int64_t)o * m + row] + ((e - e0) - before)] = r;
  }
}

}  // namespace

extern "C" int gtb_transpose_count(const int32_t* idx, int64_t nnz, int32_t col0, int32_t* cnt, int64_t n_cols,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CHECK_ARG(n_cols > 0 && nnz >= 0, "bad shape");
  GTB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n_cols, st));
  if (nnz > 0) {
    const int64_t blocks = gtb_cdiv(nnz, 256);
    transpose_count_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(idx, nnz, col0, cnt);
    GTB_CHECK_LAUNCH();
  }
  return GTB_OK;
}

extern "C" int gtb_transpose_scatter(const int64_t* indptr, const int32_t* idx, const double* val, int64_t n_rows,
                                     int32_t row0, int32_t col0, int32_t* cursor, void* t_rec, void* stream) {
  GTB_CHECK_ARG(n_rows > 0, "empty matrix");
  transpose_scatter_kernel<<<(unsigned)gtb_cdiv(n_rows * GRP, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, idx, val, n_rows, row0, col0, cursor, reinterpret_cast<EdgeRec*>(t_rec));
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_records_count(const void* rec, int64_t k, int32_t col0, int32_t* cnt, int64_t n_cols,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CHECK_ARG(n_cols > 0 && k >= 0, "bad shape");
  GTB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n_cols, st));
  if (k > 0) {
    const int64_t blocks = gtb_cdiv(k, 256);
    records_count_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const EdgeRec*>(rec), k, col0, cnt);
    GTB_CHECK_LAUNCH();
  }
  return GTB_OK;
}

extern "C" int gtb_records_scatter(const void* rec, int64_t k, int32_t col0, int32_t* cursor, void* t_rec,
                                   void* stream) {
  if (k > 0) {
    const int64_t blocks = gtb_cdiv(k, 256);
    records_scatter_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const EdgeRec*>(rec), k, col0, cursor, reinterpret_cast<EdgeRec*>(t_rec));
    GTB_CHECK_LAUNCH();
  }
  return GTB_OK;
}

extern "C" int gtb_csr_sort_rows(const int64_t* ptr, int32_t* idx, double* val, int64_t n, int32_t* has_long,
                                 void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  SoAStore st{idx, val};
  return launch_sort<SoAStore>(ptr, st, n, nullptr, 0, has_long, (cudaStream_t)stream);
}

extern "C" int gtb_rec_sort_rows(const int64_t* ptr, void* rec, int64_t n, const int64_t* pa, int min_total,
                                 int32_t* has_long, void* stream) {
  GTB_CHECK_ARG(n > 0, "empty matrix");
  RecStore st{reinterpret_cast<EdgeRec*>(rec)};
  return launch_sort<RecStore>(ptr, st, n, pa, min_total, has_long, (cudaStream_t)stream);
}

extern "C" int gtb_sym_merge_reg_rows(void) { return MRG_REG; }

extern "C" int gtb_sym_merge_count(const int64_t* pa, const int32_t* ia, const double* va, const int64_t* pt,
                                   void* t_rec, int64_t n_rows, int mode, double theta, int32_t* newlen,
                                   int32_t* worklist, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && mode >= 0 && mode <= 2, "bad arguments");
  GTB_CHECK_ARG(worklist != nullptr, "worklist: n_rows + 1 ints of scratch");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* wl_count = worklist + n_rows;
  GTB_CUDA(cudaMemsetAsync(wl_count, 0, sizeof(int32_t), st));
  sym_merge_kernel<false><<<(unsigned)gtb_cdiv(n_rows, MRG_WARPS), MRG_WARPS * 32, 0, st>>>(
      pa, ia, va, pt, reinterpret_cast<const EdgeRec*>(t_rec), n_rows, 0, mode, theta, newlen, worklist, wl_count,
      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  GTB_CHECK_LAUNCH();
  // rows too long for the register path of the fill pass: order their transposed entries by column now
  const size_t smem = (size_t)SORT_BLOCK_CAP * 12;
  GTB_CUDA(cudaFuncSetAttribute(rec_sort_worklist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = n_rows < 148 * 4 ? n_rows : 148 * 4;
  rec_sort_worklist_kernel<<<(unsigned)blocks, SORT_BLOCK_THREADS, smem, st>>>(
      pt, reinterpret_cast<EdgeRec*>(t_rec), worklist, wl_count);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_sym_merge_fill(const int64_t* pa, const int32_t* ia, const double* va, const int64_t* pt,
                                  const void* t_rec, int64_t n_rows, int32_t row0, int mode, double theta,
                                  const int64_t* outptr, int32_t* out_idx, double* out_val, double* p_val,
                                  double* degree, int32_t* flags, void* stream) {
  GTB_CHECK_ARG(n_rows > 0 && mode >= 0 && mode <= 2, "bad arguments");
  sym_merge_kernel<true><<<(unsigned)gtb_cdiv(n_rows, MRG_WARPS), MRG_WARPS * 32, 0, (cudaStream_t)stream>>>(
      pa, ia, va, pt, reinterpret_cast<const EdgeRec*>(t_rec), n_rows, row0, mode, theta, nullptr, nullptr, nullptr,
      outptr, out_idx, out_val, p_val, degree, flags);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_route_count(const int64_t* indptr, const int32_t* idx, int64_t m, int32_t per, int world,
                               int32_t* cnt, void* stream) {
  GTB_CHECK_ARG(m > 0 && per > 0 && world > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GTB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)m * world, st));
  route_count_kernel<<<(unsigned)gtb_cdiv(m * GRP, 256), 256, 0, st>>>(indptr, idx, m, per, world, cnt);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_route_fill(const int64_t* indptr, const int32_t* idx, const double* val, int64_t m, int32_t row0,
                              int32_t per, int world, const int32_t* cnt, const int64_t* pos, void* send,
                              void* stream) {
  GTB_CHECK_ARG(m > 0 && per > 0 && world > 0, "bad arguments");
  route_fill_kernel<<<(unsigned)gtb_cdiv(m * GRP, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, idx, val, m, row0, per, world, cnt, pos, reinterpret_cast<EdgeRec*>(send));
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
