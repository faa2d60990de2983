// Operand preparation for the distance/top-k stage.
//
// The search kernels consume a *search operand*: the float32 data centred on the
// reference-set column means, transposed to k-major [d_pad][n_pad] (so a K-chunk of a
// 128-point tile is a set of contiguous 512-byte rows: cp.async / TMA friendly), zero padded,
// plus squared norms of the centred rows (+inf for padding rows so they never qualify).
// Centring shrinks |x|^2 and with it the cancellation error of |x|^2+|y|^2-2xy
// (SURVEY.md H1).  The float64 re-evaluation always reads the ORIGINAL float32 rows.
#include "common.cuh"
#include "gtb200.h"

#define MEAN_BLOCKS 256

// partial[b][k] = sum over rows r == b (mod gridDim.x-strided chunks) of X[r][k]
__global__ void colsum_partial_kernel(const float* __restrict__ X, int64_t n, int d,
                                      double* __restrict__ partial) {
  // each block owns a contiguous row range; thread t sums columns t, t+blockDim, ...
  int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
  int64_t r0 = (int64_t)blockIdx.x * rows_per;
  int64_t r1 = min(n, r0 + rows_per);
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += (double)X[r * d + k];
    partial[(int64_t)blockIdx.x * d + k] = s;
  }
}

__global__ void colsum_final_kernel(const double* __restrict__ partial, int nblk, int d, int64_t n,
                                    float* __restrict__ mean) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * d + k];
  mean[k] = (float)(s / (double)n);
}

extern "C" int gtb_col_mean(const float* X, int64_t n, int d, double* ws, float* mean, void* stream) {
  GTB_CHECK_ARG(n > 0 && d > 0, "empty input");
  cudaStream_t st = (cudaStream_t)stream;
  colsum_partial_kernel<<<MEAN_BLOCKS, 128, 0, st>>>(X, n, d, ws);
  GTB_CHECK_LAUNCH();
  colsum_final_kernel<<<(d + 127) / 128, 128, 0, st>>>(ws, MEAN_BLOCKS, d, n, mean);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int64_t gtb_col_mean_ws_doubles(int d) { return (int64_t)MEAN_BLOCKS * d; }

// XT[k][r] = fl32(X[r][k] - mean[k]) for r < n, k < d; 0 elsewhere.  32x32 smem tile transpose.
__global__ void center_transpose_kernel(const float* __restrict__ X, int64_t n, int d,
                                        const float* __restrict__ mean, float* __restrict__ XT,
                                        int64_t n_pad, int d_pad) {
  __shared__ float tile[32][33];
  int64_t r0 = (int64_t)blockIdx.x * 32;
  int k0 = blockIdx.y * 32;
  // load: threadIdx.x runs along k (contiguous in X)
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i;
    int k = k0 + threadIdx.x;
    float v = 0.f;
    if (r < n && k < d) v = X[r * d + k] - (mean ? mean[k] : 0.f);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  // store: threadIdx.x runs along r (contiguous in XT)
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int k = k0 + i;
    int64_t r = r0 + threadIdx.x;
    if (k < d_pad && r < n_pad) XT[(int64_t)k * n_pad + r] = tile[threadIdx.x][i];
  }
}

// norm2[r] = sum_k XT[k][r]^2 (float64 accumulate, rounded once); +inf for r >= n.
__global__ void norms_kernel(const float* __restrict__ XT, int64_t n, int64_t n_pad, int d_pad,
                             float* __restrict__ norm2, float* __restrict__ maxnorm) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float out = 0.f;
  if (r < n_pad) {
    if (r < n) {
      double s = 0.0;
      for (int k = 0; k < d_pad; ++k) {
        double v = (double)XT[(int64_t)k * n_pad + r];
        s += v * v;
      }
      // round up so the stored norm never under-estimates (keeps the error bound one-sided safe)
      out = __double2float_ru(s);
      norm2[r] = out;
    } else {
      norm2[r] = gtb_inf_f();
    }
  }
  if (maxnorm) {
    float m = out;
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)maxnorm, __float_as_int(m));  // m >= 0
  }
}

extern "C" int gtb_prepare_operand(const float* X, int64_t n, int d, const float* mean, float* XT,
                                   int64_t n_pad, int d_pad, float* norm2, float* maxnorm,
                                   void* stream) {
  GTB_CHECK_ARG(n > 0 && d > 0 && n_pad >= n && d_pad >= d, "bad shape");
  GTB_CHECK_ARG(n_pad % 128 == 0 && d_pad % 8 == 0, "n_pad must be a multiple of 128, d_pad of 8");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(n_pad / 32), (unsigned)((d_pad + 31) / 32));
  center_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(X, n, d, mean, XT, n_pad, d_pad);
  GTB_CHECK_LAUNCH();
  if (maxnorm) GTB_CUDA(cudaMemsetAsync(maxnorm, 0, sizeof(float), st));
  norms_kernel<<<(unsigned)(n_pad / 128), 128, 0, st>>>(XT, n, n_pad, d_pad, norm2, maxnorm);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

// Gather columns of a k-major operand: dstT[k][t] = srcT[k][rows[t]], t < nt; zero padding.
__global__ void gather_operand_kernel(const float* __restrict__ srcT, int64_t src_pad,
                                      const float* __restrict__ src_n2, const int32_t* __restrict__ rows,
                                      int64_t nt, float* __restrict__ dstT, int64_t dst_pad, int d_pad,
                                      float* __restrict__ dst_n2) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= dst_pad) return;
  if (t < nt) {
    int64_t r = rows[t];
    for (int k = 0; k < d_pad; ++k) dstT[(int64_t)k * dst_pad + t] = srcT[(int64_t)k * src_pad + r];
    dst_n2[t] = src_n2[r];
  } else {
    for (int k = 0; k < d_pad; ++k) dstT[(int64_t)k * dst_pad + t] = 0.f;
    dst_n2[t] = 0.f;  // padded QUERY rows: finite norm, results discarded
  }
}

extern "C" int gtb_gather_operand(const float* srcT, int64_t src_pad, const float* src_n2,
                                  const int32_t* rows, int64_t nt, float* dstT, int64_t dst_pad,
                                  int d_pad, float* dst_n2, void* stream) {
  GTB_CHECK_ARG(nt > 0 && dst_pad >= nt && dst_pad % 128 == 0, "bad shape");
  gather_operand_kernel<<<(unsigned)gtb_cdiv(dst_pad, 128), 128, 0, (cudaStream_t)stream>>>(
      srcT, src_pad, src_n2, rows, nt, dstT, dst_pad, d_pad, dst_n2);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}
