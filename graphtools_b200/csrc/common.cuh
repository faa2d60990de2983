// Shared helpers for the gtb200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define GTB_OK 0
#define GTB_ERR_CUDA -1
#define GTB_ERR_ARG -2
#define GTB_ERR_CAPACITY -3

void gtb_set_error(const char* fmt, ...);

#define GTB_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      gtb_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, msg); \
      return GTB_ERR_ARG;                                          \
    }                                                              \
  } while (0)

#define GTB_CHECK_LAUNCH()                                         \
  do {                                                             \
    cudaError_t e_ = cudaGetLastError();                           \
    if (e_ != cudaSuccess) {                                       \
      gtb_set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return GTB_ERR_CUDA;                                         \
    }                                                              \
  } while (0)

#define GTB_CUDA(call)                                             \
  do {                                                             \
    cudaError_t e_ = (call);                                       \
    if (e_ != cudaSuccess) {                                       \
      gtb_set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return GTB_ERR_CUDA;                                         \
    }                                                              \
  } while (0)

static inline int64_t gtb_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float gtb_inf_f() { return __int_as_float(0x7f800000); }

// alpha-decay affinity exp(-(d/bw)^decay); NaN -> 1 (reference graphs.py:503-507)
// Integer decays up to 64 (the default is 40) are evaluated by binary exponentiation: at most 6 squarings and 5
// products instead of the ~150 instructions of the general double-precision pow -- within 6 ulp of it, i.e. a relative
// difference below 1e-14 in the affinity at the threshold (where (d/bw)^decay = 9.2), far inside the 1e-5 tolerance.
__device__ __forceinline__ double gtb_pow_decay(double x, double decay) {
  const int n = (int)decay;
  if ((double)n == decay && n >= 1 && n <= 64) {
    double r = (n & 1) ? x : 1.0, b = x;
#pragma unroll
    for (int bit = 1; bit < 7; ++bit) {
      b *= b;
      if ((n >> bit) & 1) r *= b;
    }
    return r;
  }
  return pow(x, decay);
}
__device__ __forceinline__ double gtb_affinity(double dist, double bw, double decay) {
  double w = exp(-gtb_pow_decay(dist / bw, decay));
  return (w != w) ? 1.0 : w;
}

// Bitonic sort of 32 (key, value) pairs held one per lane, ascending by (key, value), entirely in registers: fifteen
// compare-exchange steps of three shuffles each -- no shared memory, no barriers, no divergent branches (the
// comparisons are combined with bitwise operators so that ptxas emits predicates, not BSSY/BRA regions).
template <typename K, typename V>
__device__ __forceinline__ void warp_sort32(K& k, V& v, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const K ok = __shfl_xor_sync(0xffffffffu, k, j);
      const V ov = __shfl_xor_sync(0xffffffffu, v, j);
      const bool keep_min = (((lane & kk) == 0) == ((lane & j) == 0));
      const bool gt = (k > ok) | ((k == ok) & (v > ov));
      const bool lt = (k < ok) | ((k == ok) & (v < ov));
      const bool take = keep_min ? gt : lt;
      k = take ? ok : k;
      v = take ? ov : v;
    }
  }
}

// Bitonic sort of n_pow2 (key, payload) pairs living in shared memory, ascending by
// (key, idx).  Executed by `nthreads` cooperating threads whose rank is `tid`;
// SYNC() must be a barrier over exactly those threads.
#define GTB_BITONIC_SORT(KEY, IDX, n_pow2, tid, nthreads, SYNC, KEY_T, IDX_T)        \
  for (int k_ = 2; k_ <= (n_pow2); k_ <<= 1) {                                     \
    for (int j_ = k_ >> 1; j_ > 0; j_ >>= 1) {                                     \
      for (int t_ = (tid); t_ < (n_pow2); t_ += (nthreads)) {                      \
        int p_ = t_ ^ j_;                                                          \
        if (p_ > t_) {                                                             \
          KEY_T a_ = KEY[t_], b_ = KEY[p_];                                        \
          IDX_T ia_ = IDX[t_], ib_ = IDX[p_];                                      \
          bool up_ = ((t_ & k_) == 0);                                             \
          bool gt_ = (a_ > b_) || (a_ == b_ && ia_ > ib_);                         \
          if (gt_ == up_) {                                                        \
            KEY[t_] = b_; KEY[p_] = a_; IDX[t_] = ib_; IDX[p_] = ia_;              \
          }                                                                        \
        }                                                                          \
      }                                                                            \
      SYNC();                                                                      \
    }                                                                              \
  }
