// Batched GEMM C[b] = alpha * A[b] . B[b]^T on tcgen05 / TMEM / TMA with fp32-class accuracy (sm_100a): the landmark
// contractions of NystromAttention (models/NystromAttention.py:89,122-125,138-140,150 - to_qkv, the three similarity
// products, the 24 products of the 6-step pseudo-inverse per head, (attn1 @ attn2_inv) @ (attn3 @ v), to_out) and
// their gradients.
//
// The tensor cores take 16-bit operands, and 11 bits do not survive the pseudo-inverse recurrence at 1e-3 (TF32: 6e-3,
// DESIGN.md section 3).  Every fp32 operand x is therefore split once into two fp16 tensors, x * s = hi + lo with a
// power-of-two scale s that puts max|x| at 2^10 (dml_split_f16: 22 significant bits, no underflow of the small softmax
// entries), and the product is accumulated as hi.hi + hi.lo + lo.hi in fp32 in TMEM (three tcgen05.mma per k-step);
// alpha = 1 / (s_A s_B) (times the caller's factor) is applied in the epilogue.  dml_split_f16 can transpose while it
// splits, so every product in the forward and backward is brought to the one layout the kernel implements: A [M, K] and
// B [N, K], both K-contiguous (the layout of S = Q K^T).
//
// CTA = one 128 x BN output tile (BN = 128 or 64) of one batch entry; warp 4 = TMA producer (3-stage ring of
// {A_hi, A_lo, B_hi, B_lo} 64-wide k-blocks, 128-byte swizzle), warp 5 = MMA issuer + TMEM owner, warps 0-3 = epilogue
// (TMEM lane = output row).
#include <math.h>

#include "../../include/dml_b200.h"
#include "tc_common.cuh"

namespace dml {
namespace tc {
namespace gemm {

constexpr int kBM = 128, kBK = 64, kStages = 3, kThreads = 32 * 6;
constexpr uint32_t kTileA = kBM * kBK * 2;   // 16 KB

template <int BN>
struct Cfg {
  static constexpr uint32_t kTileB = BN * kBK * 2;
  static constexpr uint32_t kStageBytes = 2 * kTileA + 2 * kTileB;
  static constexpr uint32_t kOffBar = kStages * kStageBytes;
  static constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAcc = 2 * kStages, kNumBars = 2 * kStages + 1;
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kSmemBytes = kOffTmemPtr + 16 + 1024;
  static constexpr uint32_t kIdesc = idesc_f16(128, BN, false, false);
};

struct Params {
  float* c;                 // [batch, M, ldc]
  const float* scale_a;     // device float[2] = (s, 1/s) of A (dml_split_f16)
  const float* scale_b;
  float alpha;
  long long c_batch_stride; // elements
  int M, N, K, ldc;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_nt_split_kernel(const __grid_constant__ CUtensorMap mah, const __grid_constant__ CUtensorMap mal,
                     const __grid_constant__ CUtensorMap mbh, const __grid_constant__ CUtensorMap mbl, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN, b = blockIdx.z;
  const int nk = cdiv(p.K, kBK);
  auto bar = [&](int i) { return sbase + C::kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + C::kOffTmemPtr);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(C::kBarFull + s), 1); mbar_init(bar(C::kBarEmpty + s), 1); }
    mbar_init(bar(C::kBarAcc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::kOffTmemPtr), "r"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---- TMA producer ----
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % kStages;
        mbar_wait(bar(C::kBarEmpty + st), ((kb / kStages) & 1) ^ 1);
        const uint32_t dst = sbase + st * C::kStageBytes;
        mbar_expect_tx(bar(C::kBarFull + st), C::kStageBytes);
        tma_load_3d(dst, &mah, bar(C::kBarFull + st), kb * kBK, m0, b);
        tma_load_3d(dst + kTileA, &mal, bar(C::kBarFull + st), kb * kBK, m0, b);
        tma_load_3d(dst + 2 * kTileA, &mbh, bar(C::kBarFull + st), kb * kBK, n0, b);
        tma_load_3d(dst + 2 * kTileA + C::kTileB, &mbl, bar(C::kBarFull + st), kb * kBK, n0, b);
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer (uniform datapath, one elected lane) ----
    const bool leader = elect_one();
    for (int kb = 0; kb < nk; ++kb) {
      const int st = kb % kStages;
      mbar_wait(bar(C::kBarFull + st), (kb / kStages) & 1);
      tc_fence_after();
      const uint32_t base = sbase + st * C::kStageBytes;
      const uint64_t ah = smem_desc(base), al = smem_desc(base + kTileA);
      const uint64_t bh = smem_desc(base + 2 * kTileA), bl = smem_desc(base + 2 * kTileA + C::kTileB);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) {
        mma_ss(tmem, ah + 2 * k, bh + 2 * k, C::kIdesc, (kb > 0) || (k > 0), leader);
        mma_ss(tmem, ah + 2 * k, bl + 2 * k, C::kIdesc, 1, leader);
        mma_ss(tmem, al + 2 * k, bh + 2 * k, C::kIdesc, 1, leader);
      }
      tc_commit(bar(C::kBarEmpty + st), leader);
    }
    tc_commit(bar(C::kBarAcc), leader);
  } else {
    // ---- epilogue: TMEM lane = output row ----
    mbar_wait(bar(C::kBarAcc), 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const float alpha = p.alpha * __ldg(p.scale_a + 1) * __ldg(p.scale_b + 1);
    float* crow = p.c + (size_t)b * p.c_batch_stride + (size_t)row * p.ldc + n0;
    const uint32_t tb = tmem + (((uint32_t)warp * 32u) << 16);
    const bool vec = ((p.ldc & 3) == 0) && ((((uintptr_t)p.c) & 15) == 0) && ((p.c_batch_stride & 3) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t a[16];
      tmem_ld16(tb + c * 16, a);
      tmem_ld_wait(a);
      if (row < p.M) {
        const int col0 = n0 + c * 16;
        if (vec && col0 + 16 <= p.N) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(crow + c * 16 + e) =
                make_float4(__uint_as_float(a[e]) * alpha, __uint_as_float(a[e + 1]) * alpha,
                            __uint_as_float(a[e + 2]) * alpha, __uint_as_float(a[e + 3]) * alpha);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (col0 + e < p.N) crow[c * 16 + e] = __uint_as_float(a[e]) * alpha;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN));
  }
}

// max |x| over a strided [batch, R, C] tensor -> atomicMax on the bit pattern (non-negative floats order like uints)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, long long bs, int batch, int R, int C, int ld, uint32_t* __restrict__ out) {
  float m = 0.f;
  const long long per = (long long)R * C, total = per * batch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const long long rc = i - (long long)b * per;
    const int r = (int)(rc / C), c = (int)(rc - (long long)r * C);
    m = fmaxf(m, fabsf(x[(size_t)b * bs + (size_t)r * ld + c]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// power-of-two scale that puts max|x| into [2^9, 2^10)  (1 for an all-zero tensor)
__device__ __forceinline__ float split_scale(const uint32_t* amax_bits) {
  const float amax = __uint_as_float(__ldg(amax_bits));
  if (!(amax > 0.f) || !isfinite(amax)) return 1.0f;
  return exp2f(fminf(fmaxf(floorf(log2f(1024.0f / amax)), -100.f), 100.f));
}

// x [batch, R, C] fp32 (row stride ld, batch stride bs) -> hi, lo fp16 [batch, R, ldo] (transpose = 0) or [batch, C, ldo]
// (transpose = 1), hi + lo = x * s; columns past the logical width of the padded leading dimension are zero-filled.
// Block (0,0,0) publishes scale[0..1] = (s, 1/s).
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, long long bs, int R, int C, int ld, const uint32_t* __restrict__ amax_bits,
                 float* __restrict__ scale, int transpose, int ldo, h16* __restrict__ hi, h16* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float s = split_scale(amax_bits);
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) { scale[0] = s; scale[1] = 1.0f / s; }
  const float* xb = x + (size_t)b * bs;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (!transpose) {
    const size_t ob = (size_t)b * R * ldo;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      if (r < R && c < ldo) {
        const float v = c < C ? xb[(size_t)r * ld + c] * s : 0.f;
        const h16 h = __float2half_rn(v);
        hi[ob + (size_t)r * ldo + c] = h;
        lo[ob + (size_t)r * ldo + c] = __float2half_rn(v - __half2float(h));
      }
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < R && c < C) ? xb[(size_t)r * ld + c] * s : 0.f;
    }
    __syncthreads();
    const size_t ob = (size_t)b * C * ldo;
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;      // output row = source column
      if (c < C && r < ldo) {
        const float v = tile[tx][i];
        const h16 h = __float2half_rn(v);
        hi[ob + (size_t)c * ldo + r] = h;
        lo[ob + (size_t)c * ldo + r] = __float2half_rn(v - __half2float(h));
      }
    }
  }
}

}  // namespace gemm
}  // namespace tc
}  // namespace dml

extern "C" {

int dml_split_f16(const float* x, long long batch_stride, int batch, int R, int C, int ld, int transpose, int ldo,
                  void* hi, void* lo, float* scale, void* amax_ws, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && scale && amax_ws && hi && lo && batch > 0 && R > 0 && C > 0 && ld >= C && (ldo % 8) == 0);
  DML_CHECK_ARG(ldo >= (transpose ? R : C));
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(amax_ws, 0, 4, st);
  if (e != cudaSuccess) return (int)e;
  const long long total = (long long)batch * R * C;
  const int blocks = (int)min((total + 255) / 256, (long long)148 * 8);
  tc::gemm::absmax_kernel<<<blocks, 256, 0, st>>>(x, batch_stride, batch, R, C, ld, (uint32_t*)amax_ws);
  dim3 grid;
  if (!transpose) grid = dim3(cdiv(ldo, 32), cdiv(R, 32), batch);
  else grid = dim3(cdiv(C, 32), cdiv(ldo, 32), batch);
  tc::gemm::split_f16_kernel<<<grid, 256, 0, st>>>(x, batch_stride, R, C, ld, (const uint32_t*)amax_ws, scale, transpose,
                                                  ldo, (h16*)hi, (h16*)lo);
  DML_RETURN_LAUNCH();
}

/* C[b] (float [M, ldc], batch stride c_batch_stride) = alpha / (s_A s_B) * A[b] . B[b]^T,  A = a_hi + a_lo [batch, M, lda],
 * B = b_hi + b_lo [batch, N, ldb] (fp16, K-contiguous, lda / ldb multiples of 8, K <= lda, ldb).                       */
int dml_gemm_nt_split(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* scale_a,
                      const float* scale_b, float alpha, int batch, int M, int N, int K, int lda, int ldb, float* c,
                      int ldc, long long c_batch_stride, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  using namespace dml::tc::gemm;
  DML_CHECK_ARG(a_hi && a_lo && b_hi && b_lo && scale_a && scale_b && c && batch > 0 && M > 0 && N > 0 && K > 0);
  if ((lda % 8) || (ldb % 8) || lda < K || ldb < K || ldc < N) return DML_EINVAL;
  if ((((uintptr_t)a_hi) | ((uintptr_t)a_lo) | ((uintptr_t)b_hi) | ((uintptr_t)b_lo)) & 15) return DML_EINVAL;
  const int BN = N > 64 ? 128 : 64;
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  // the K extent of the maps is the logical K: columns K..ld-1 and rows past M / N are read as zero
  auto mk = [&](CUtensorMap* m, const void* base, int rows, int ld, int box_rows) -> int {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return DML_EUNSUPPORTED;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? DML_OK : DML_EINVAL;
  };
  if ((rc = mk(&mah, a_hi, M, lda, kBM)) || (rc = mk(&mal, a_lo, M, lda, kBM)) || (rc = mk(&mbh, b_hi, N, ldb, BN)) ||
      (rc = mk(&mbl, b_lo, N, ldb, BN)))
    return rc;
  Params p{};
  p.c = c; p.scale_a = scale_a; p.scale_b = scale_b; p.alpha = alpha; p.c_batch_stride = c_batch_stride;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_split_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_nt_split_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64>::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(cdiv(M, kBM), cdiv(N, BN), batch);
  if (BN == 128)
    gemm_nt_split_kernel<128><<<grid, kThreads, Cfg<128>::kSmemBytes, (cudaStream_t)stream>>>(mah, mal, mbh, mbl, p);
  else
    gemm_nt_split_kernel<64><<<grid, kThreads, Cfg<64>::kSmemBytes, (cudaStream_t)stream>>>(mah, mal, mbh, mbl, p);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
