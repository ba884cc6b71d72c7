// Shared device helpers for the dml_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "dml_b200 kernels are written for sm_100a (Blackwell B200) only"
#endif

#include "../../include/dml_b200.h"

#define DML_OK 0
#define DML_EINVAL DML_E_INVAL
#define DML_EUNSUPPORTED DML_E_UNSUPPORTED
#define DML_EWORKSPACE DML_E_WORKSPACE

#define DML_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return DML_EINVAL; \
  } while (0)

// Return the CUDA error of the launch that just happened (no sync).
#define DML_RETURN_LAUNCH()                    \
  do {                                         \
    cudaError_t e__ = cudaGetLastError();      \
    return e__ == cudaSuccess ? DML_OK : (int)e__; \
  } while (0)

namespace dml {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;
typedef __half h16;      // 16-bit operand type of the attention core (fp16: 11-bit significand, fp32 accumulate)

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// (x0, x1) -> packed fp16 pair `hi` and the packed fp16 residual pair `lo` (x ~= hi + lo to 22 bits)
// The residual x - hi is one mixed-precision FMA per element (fma.rn.f32.f16: fp16 x fp16 + fp32, exact here) instead of an
// unpack and a subtract.
__device__ __forceinline__ void split_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  float r0, r1;
  asm("{\n.reg .b16 h0, h1, m1;\n"
      "cvt.rn.f16x2.f32 %0, %4, %3;\n"
      "mov.b32 {h0, h1}, %0;\n"
      "mov.b16 m1, 0xBC00;\n"                      // -1.0
      "fma.rn.f32.f16 %1, h0, m1, %3;\n"
      "fma.rn.f32.f16 %2, h1, m1, %4;\n}"
      : "=&r"(hi), "=f"(r0), "=f"(r1) : "f"(x0), "f"(x1));
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(lo) : "f"(r0), "f"(r1));
}

// ---- bf16 operand pairs (x ~= hi + lo, 16 significant bits, fp32 exponent range: no scale) -------------------------
__device__ __forceinline__ float bf16_lo_f(uint32_t w) { return __uint_as_float(w << 16); }          // low half of a packed pair
__device__ __forceinline__ float bf16_hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }  // high half
// (x0, x1) -> packed bf16 pair `hi` and packed bf16 residual pair `lo`
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float r0 = x0 - bf16_lo_f(hi), r1 = x1 - bf16_hi_f(hi);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// four consecutive values -> 8 bytes into each plane
__device__ __forceinline__ void store_pair4(bf16* hi_ptr, bf16* lo_ptr, float a, float b, float c, float d) {
  uint32_t h0, l0, h1, l1;
  split_bf16x2(a, b, h0, l0);
  split_bf16x2(c, d, h1, l1);
  *reinterpret_cast<uint2*>(hi_ptr) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo_ptr) = make_uint2(l0, l1);
}
// four consecutive values of a pair (8-byte aligned) -> fp32
__device__ __forceinline__ float4 load_pair4(const bf16* hi_ptr, const bf16* lo_ptr) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi_ptr), l = *reinterpret_cast<const uint2*>(lo_ptr);
  return make_float4(bf16_lo_f(h.x) + bf16_lo_f(l.x), bf16_hi_f(h.x) + bf16_hi_f(l.x), bf16_lo_f(h.y) + bf16_lo_f(l.y),
                     bf16_hi_f(h.y) + bf16_hi_f(l.y));
}

// ---- cp.async (LDGSTS) 16-byte copies with zero-fill predicate ----
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool pred) {
  int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- ldmatrix / mma.sync (legacy warp-level tensor path, used by the v1 attention kernels) ----
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tile of [rows][64] 16-bit elements (128 B per row = 8 chunks of 16 B), XOR-swizzled on the chunk index so
// that ldmatrix (8 rows x 16 B) and 16-B row-wise stores are bank-conflict free.
__device__ __forceinline__ int swz64(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

}  // namespace dml
