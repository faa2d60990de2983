// K9: float64-faithful dense GEMM on the INT8 tensor cores (tcgen05.mma kind::i8), the building block of the
// landmark diffusion chain landmark_op^t (the dense L x L products the callers of G.landmark_op run
// through np.linalg.matrix_power; operator built at reference graphtools/graphs.py:1240-1243).
//
// Digit slicing (Ozaki-style): every row i of the left operand is scaled by a power of two sa_i so that |a| < 1/4 and
// written as S signed base-256 digits, a = sum_s d_s 2^(-8(s+1)), d_s in [-128, 127] (balanced, exact: the digits ARE
// the fixed-point number); the right operand likewise per column.  Then
//     C_ij = sa_i sb_j sum_{o = 0}^{S-1} 2^(-8(o+2)) * ACC_o[i][j],     ACC_o = sum_{s + t = o} sum_k dA_s[i,k] dB_t[j,k]
// and every ACC_o is an EXACT int32 sum on the tensor cores (|d d'| <= 2^14, K <= 16384, at most 8 digit pairs per
// order: < 2^31).  The only error is the fixed-point truncation of the inputs at 2^(-8S) of their row / column maximum
// and the dropped digit pairs of order >= S: |err| <= (S + 2) K 2^(-8S - 2) sa_i sb_j -- with S = 7 below the rounding
// of a float64 dot product of the same length; the final sum over o is one float64 rounding per term.  The result does
// not depend on tiling, accumulation order or rank count.
//
// Kernel: one CTA per 128 x 64 output tile; ALL S accumulators (S x 64 TMEM columns of int32) stay resident, so each
// k-block of the S digit planes of A (128 rows) and B (64 rows) is loaded ONCE by TMA and used for all S(S+1)/2 digit
// pairs.  Warp 0 = TMA producer, warp 1 = MMA issuer (SS mode, K-major operands, swizzled rows of RB bytes), warps 2-5
// = epilogue (tcgen05.ld 32x32b.x16, int32 -> float64, scale, store).
#include "common.cuh"
#include "gtb200.h"
#include "tc_ptx.cuh"

namespace {
using namespace gtbptx;

constexpr int GM = 128, GN = 64, G_THREADS = 64 + 128;

struct GemmParams {
  int64_t M, N, M_pad, N_pad;
  int nkb;                       // k-blocks of RB bytes
  const double* sa; const double* sb;
  double* C; int64_t ldc;
};

__device__ __forceinline__ void mma_ss_i8_pred(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n .reg .pred p, q;\n setp.ne.b32 p, %4, 0;\n setp.ne.b32 q, %5, 0;\n"
      " @q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void commit_pred(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %1, 0;\n"
      " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"
      ::"r"(bar), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// S digit planes, rows of RB bytes (= RB int8 elements along K) per stage: RB = 128 -> 1 stage, 64 -> 2, 32 -> 4
template <int S, int RB>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - raw);
  constexpr int NSTAGE = 128 / RB;
  constexpr uint32_t A_BYTES = GM * RB, B_BYTES = GN * RB, STAGE_BYTES = S * (A_BYTES + B_BYTES);
  constexpr uint32_t LAYOUT = (RB == 128) ? 2u : (RB == 64 ? 4u : 6u);
  constexpr int KSTEPS = RB / 32;
  static_assert(S >= 2 && S * GN <= 512, "all S accumulators must fit the 512 TMEM columns");
  const uint32_t bar0 = base + NSTAGE * STAGE_BYTES;
  const uint32_t full_b = bar0, empty_b = bar0 + 32, tm_full = bar0 + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + (bar0 - base) + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  const int nkb = p.nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_b + 8 * s, 1);
      mbar_init(empty_b + 8 * s, 1);
    }
    mbar_init(tm_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: all digit planes of one k-block per stage =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NSTAGE;
        const uint32_t ph = (uint32_t)((kb / NSTAGE) & 1);
        mbar_wait(empty_b + 8 * s, ph ^ 1);
        mbar_arrive_expect_tx(full_b + 8 * s, STAGE_BYTES);
        const uint32_t dst = base + s * STAGE_BYTES;
        for (int sl = 0; sl < S; ++sl)
          tma_load_2d(dst + sl * A_BYTES, &mapA, full_b + 8 * s, kb * RB, (int)(sl * p.M_pad + m0));
        for (int sl = 0; sl < S; ++sl)
          tma_load_2d(dst + S * A_BYTES + sl * B_BYTES, &mapB, full_b + 8 * s, kb * RB, (int)(sl * p.N_pad + n0));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow, instructions predicated) ============
    const uint32_t lead = elect_one() ? 1u : 0u;
    // instruction descriptor: D = S32 (2), A = B = S8 (1), K-major, N = 64, M = 128
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GN >> 3) << 17) |
                               ((uint32_t)(GM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % NSTAGE;
      const uint32_t ph = (uint32_t)((kb / NSTAGE) & 1);
      mbar_wait(full_b + 8 * s, ph);
      tc_fence_after();
      const uint32_t st = base + s * STAGE_BYTES;
      const uint64_t ad0 = make_desc(st, 8 * RB, LAYOUT);
      const uint64_t bd0 = make_desc(st + S * A_BYTES, 8 * RB, LAYOUT);
#pragma unroll 1
      for (int da = 0; da < S; ++da) {
        const uint64_t ad = ad0 + (uint64_t)((da * A_BYTES) >> 4);
#pragma unroll 1
        for (int db = 0; db + da < S; ++db) {
          const uint64_t bd = bd0 + (uint64_t)((db * B_BYTES) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)((da + db) * GN);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk)
            mma_ss_i8_pred(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb | da | kk) != 0 ? 1u : 0u, lead);
        }
      }
      commit_pred(empty_b + 8 * s, lead);          // the stage is free once these MMAs retire
    }
    commit_pred(tm_full, lead);                      // every accumulator is complete
  } else {
    // ===================== epilogue: thread == output row, 16 columns of all S accumulators at a time ==========
    const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int64_t gi = m0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    mbar_wait(tm_full, 0);
    tc_fence_after();
    const double srow = (gi < p.M) ? p.sa[gi] : 0.0;
#pragma unroll 1
    for (int cc = 0; cc < GN / 16; ++cc) {
      uint32_t r[S][16];
      __syncwarp();
#pragma unroll
      for (int o = 0; o < S; ++o) tmem_ld16_nowait(lane_addr + (uint32_t)(o * GN + cc * 16), r[o]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int o = 0; o < S; ++o) {
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("" : "+r"(r[o][j]));
      }
      double acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0;
      // smallest order first: each term is an exact integer times a power of two, one rounding per addition
#pragma unroll
      for (int o = S - 1; o >= 0; --o) {
        const double w = __longlong_as_double((long long)(1023 - 8 * (o + 2)) << 52);   // 2^(-8 (o + 2))
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma((double)(int)r[o][j], w, acc[j]);
      }
      if (gi < p.M) {
        const int64_t c0 = n0 + cc * 16;
        double* out = p.C + gi * p.ldc + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.N) out[j] = acc[j] * srow * p.sb[c0 + j];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- digit slicing
// scale[r] = 2^(e + 2) with max_k |row r| < 2^e (1 for an all-zero row): |v / scale| < 1/4
__device__ __forceinline__ double scale_of_max(double mx) {
  if (!(mx > 0.0)) return 1.0;
  int e;
  frexp(mx, &e);                                   // mx = m 2^e, m in [0.5, 1)
  return ldexp(1.0, e + 2);
}

__global__ void slice_rowmax_kernel(const double* __restrict__ X, int64_t R, int64_t K, int64_t ld,
                                    double* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  double mx = 0.0;
  for (int64_t k = lane; k < K; k += 32) mx = fmax(mx, fabs(X[r * ld + k]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) scale[r] = scale_of_max(mx);
}

// transposed operand: logical row r = column r of X[K][ld]
__global__ void slice_colmax_kernel(const double* __restrict__ X, int64_t R, int64_t K, int64_t ld,
                                    double* __restrict__ scale) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double mx = 0.0;
  for (int64_t k = 0; k < K; ++k) mx = fmax(mx, fabs(X[k * ld + r]));
  scale[r] = scale_of_max(mx);
}

// digits[s][r][k] (int8, [S][R_pad][K_pad], zero padded); block = 32 x 8 threads, tile = 32 rows x 32 k
template <bool T>
__global__ void slice_digits_kernel(const double* __restrict__ X, int64_t R, int64_t K, int64_t ld,
                                    const double* __restrict__ scale, int S, int64_t R_pad, int64_t K_pad,
                                    int8_t* __restrict__ dig) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
  if (T) {
    for (int i = ty; i < 32; i += 8) {
      const int64_t k = k0 + i, r = r0 + tx;
      tile[i][tx] = (k < K && r < R) ? X[k * ld + r] : 0.0;
    }
    __syncthreads();
  }
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, k = k0 + tx;
    double v;
    if (T) v = tile[tx][i];
    else v = (r < R && k < K) ? X[r * ld + k] : 0.0;
    long long Q = 0;
    if (r < R && v != 0.0) Q = __double2ll_rn(scalbn(v / scale[r], 8 * S));
    for (int s = S - 1; s >= 0; --s) {
      const int8_t d = (int8_t)(Q & 0xff);                       // balanced digit in [-128, 127]
      Q = (Q - (long long)d) >> 8;
      dig[((int64_t)s * R_pad + r) * K_pad + k] = d;
    }
  }
}

int make_map_u8(CUtensorMap* m, const void* ptr, int64_t rows, int64_t K_pad, int box_k, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { gtb_set_error("cuTensorMapEncodeTiled entry point not available"); return GTB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K_pad};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_k == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { gtb_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GTB_ERR_CUDA; }
  return GTB_OK;
}

template <int S, int RB>
int launch_gemm(const int8_t* Ad, const int8_t* Bd, GemmParams& p, int64_t K_pad, cudaStream_t st) {
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = make_map_u8(&mapA, Ad, S * p.M_pad, K_pad, RB, GM))) return rc;
  if ((rc = make_map_u8(&mapB, Bd, S * p.N_pad, K_pad, RB, GN))) return rc;
  p.nkb = (int)(K_pad / RB);
  const size_t smem = 1024 + (size_t)(128 / RB) * S * (GM + GN) * RB + 128;
  auto kern = gemm_i8_kernel<S, RB>;
  GTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(p.N_pad / GN), (unsigned)(p.M_pad / GM));
  kern<<<grid, G_THREADS, smem, st>>>(mapA, mapB, p);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

template <int S>
int launch_gemm_rb(const int8_t* Ad, const int8_t* Bd, GemmParams& p, int64_t K_pad, int rb, cudaStream_t st) {
  if (rb == 128) return launch_gemm<S, 128>(Ad, Bd, p, K_pad, st);
  if (rb == 64) return launch_gemm<S, 64>(Ad, Bd, p, K_pad, st);
  return launch_gemm<S, 32>(Ad, Bd, p, K_pad, st);
}

}  // namespace

extern "C" int gtb_gemm_max_k(void) { return 16384; }

extern "C" int gtb_slice_f64(const double* X, int64_t R, int64_t K, int64_t ld, int transposed, int slices,
                             int64_t R_pad, int64_t K_pad, int8_t* digits, double* scale, void* stream) {
  GTB_CHECK_ARG(R > 0 && K > 0 && R_pad >= R && K_pad >= K, "bad shape");
  GTB_CHECK_ARG(slices >= 2 && slices <= 7, "2 to 7 digit planes");
  GTB_CHECK_ARG(R_pad % 128 == 0 && K_pad % 128 == 0, "pads must be multiples of 128");
  GTB_CHECK_ARG(ld >= (transposed ? R : K), "leading dimension too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (transposed) slice_colmax_kernel<<<(unsigned)gtb_cdiv(R, 128), 128, 0, st>>>(X, R, K, ld, scale);
  else slice_rowmax_kernel<<<(unsigned)gtb_cdiv(R * 32, 256), 256, 0, st>>>(X, R, K, ld, scale);
  GTB_CHECK_LAUNCH();
  dim3 grid((unsigned)(K_pad / 32), (unsigned)(R_pad / 32)), block(32, 8);
  if (transposed) slice_digits_kernel<true><<<grid, block, 0, st>>>(X, R, K, ld, scale, slices, R_pad, K_pad, digits);
  else slice_digits_kernel<false><<<grid, block, 0, st>>>(X, R, K, ld, scale, slices, R_pad, K_pad, digits);
  GTB_CHECK_LAUNCH();
  return GTB_OK;
}

extern "C" int gtb_gemm_i8(const int8_t* a_digits, const int8_t* b_digits, int slices, int64_t M, int64_t N,
                           int64_t K_pad, int64_t M_pad, int64_t N_pad, const double* sa, const double* sb, double* C,
                           int64_t ldc, int row_bytes, void* stream) {
  GTB_CHECK_ARG(M > 0 && N > 0 && M_pad >= M && N_pad >= N && ldc >= N, "bad shape");
  GTB_CHECK_ARG(M_pad % 128 == 0 && N_pad % 128 == 0 && K_pad % 128 == 0 && K_pad > 0, "pads must be multiples of 128");
  GTB_CHECK_ARG(K_pad <= 16384, "K too large for exact int32 accumulation (gtb_gemm_max_k)");
  GTB_CHECK_ARG(slices >= 2 && slices <= 7, "2 to 7 digit planes");
  GTB_CHECK_ARG(row_bytes == 32 || row_bytes == 64 || row_bytes == 128, "row_bytes must be 32, 64 or 128");
  GTB_CHECK_ARG((int64_t)slices * M_pad < (1ll << 31) && (int64_t)slices * N_pad < (1ll << 31), "too many rows");
  GemmParams p{};
  p.M = M; p.N = N; p.M_pad = M_pad; p.N_pad = N_pad; p.sa = sa; p.sb = sb; p.C = C; p.ldc = ldc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (slices) {
    case 2: return launch_gemm_rb<2>(a_digits, b_digits, p, K_pad, row_bytes, st);
    case 3: return launch_gemm_rb<3>(a_digits, b_digits, p, K_pad, row_bytes, st);
    case 4: return launch_gemm_rb<4>(a_digits, b_digits, p, K_pad, row_bytes, st);
    case 5: return launch_gemm_rb<5>(a_digits, b_digits, p, K_pad, row_bytes, st);
    case 6: return launch_gemm_rb<6>(a_digits, b_digits, p, K_pad, row_bytes, st);
    default: return launch_gemm_rb<7>(a_digits, b_digits, p, K_pad, row_bytes, st);
  }
}
