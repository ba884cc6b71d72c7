// Shared device helpers of the tcgen05 / TMEM / TMA attention kernels (sm_100a): mbarrier, TMA, tcgen05.mma/ld/st
// wrappers, shared-memory matrix descriptors, instruction descriptors, host-side tensor maps.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "cpb_table.cuh"

namespace dml {
namespace tc {

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {   // no arrival
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"      // %2: suspend-time hint, so a waiting warp
      "@P1 bra DONE;\n"                                                     // sleeps in the barrier unit instead of
      "bra LAB_WAIT;\n"                                                     // re-polling through the issue slots
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// the same wait for warps that are usually early (elementwise warps waiting for the next tile): back off between polls so
// the polling does not take issue slots from the warps of the same scheduler that still have work
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "LAB_WAIT:\n"
      "nanosleep.u32 40;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The MMA-issuing warp runs its loop with all 32 lanes under provably warp-uniform control flow (warp index obtained
// with a shuffle, see warp_index_uniform), so operands stay in uniform registers; `leader` (one elected lane, chosen
// once) predicates each tcgen05.mma / tcgen05.commit.
__device__ __forceinline__ int warp_index_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\nselp.u32 %0, 1, 0, e;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar, bool leader) {
  asm volatile(
      "{\n.reg .pred e;\nsetp.ne.b32 e, %1, 0;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar), "r"((uint32_t)leader) : "memory");
}
// D[tmem] (+)= A[smem] B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, bool leader) {
  asm volatile(
      "{\n.reg .pred p, e;\nsetp.ne.b32 e, %5, 0;\nsetp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"((uint32_t)leader) : "memory");
}
// D[tmem] (+)= A[tmem] B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc, bool leader) {
  asm volatile(
      "{\n.reg .pred p, e;\nsetp.ne.b32 e, %5, 0;\nsetp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"((uint32_t)leader) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for the TMEM loads; the registers are listed so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait2(uint32_t (&a)[16], uint32_t (&b)[16]) {
  tmem_ld_wait(a);
  asm volatile("" : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
               "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared-memory matrix descriptor, 128-byte swizzle, 8-row groups 1024 B apart (K-major and MN-major alike here)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor of kind::f16: fp16 x fp16 -> fp32, M x N tile; a_mn / b_mn = operand is MN-major in shared memory
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float seq_pos(int i, int n) { return (2.0f * (float)i) / (float)max(n - 1, 1) - 1.0f; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// make the results of every outstanding tcgen05.ld of this thread usable: lists the destination registers
__device__ __forceinline__ void tmem_ld_fence() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void reg_fence(uint32_t (&r)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}

// ---- compact shared-memory image of the bias table (both head outputs) --------------------------------------
//   bp    float[kCpbCells]        first breakpoint inside the cell (x units; +inf none, NaN: >= 2 breakpoints)
//   piece float4[kCpbCells][2]    (a0, c0, a1, c1) left / right of that breakpoint
//   meta  uint32[kCpbCells + 1]   low 16 bits: segment index at the start of the cell; high 16 bits: number of
//                                 flagged (NaN) cells before this one
constexpr uint32_t kTabSmemBytes = kCpbCells * 4 + kCpbCells * 32 + (kCpbCells + 1) * 4 + 12;   // multiple of 16
struct Lookup {
  uint32_t bp, piece, meta;   // shared-space addresses
  const uint32_t* gtab;       // table in global memory (slow path)
  float c1, c2;               // cell = floor(x c1 + c2)
  float c2m;                  // c2 - 0.5: cell_index() rounds on the FMA pipe instead of converting on the XU pipe
  uint32_t bp_adj, piece_adj; // bp - 4 K, piece - 32 K (mod 2^32), K = 0x4B400000: the raw rounded float of cell_raw() scales
                              // straight into an address, without first subtracting K
};
// floor(x c1 + c2) for 0 <= value < 2^22 without a float->int conversion: adding 1.5 * 2^23 leaves the integer part in the
// low mantissa bits (round-to-nearest of value - 0.5 == floor(value) except on exact ties, where either neighbouring cell
// describes the same continuous function).
__device__ __forceinline__ uint32_t cell_raw(const Lookup& L, float x) {      // K + cell index
  return __float_as_uint(fmaf(x, L.c1, L.c2m) + 12582912.0f);
}
__device__ __forceinline__ int cell_index(const Lookup& L, float x) {
  return __float_as_int(fmaf(x, L.c1, L.c2m) + 12582912.0f) - 0x4B400000;      // 12582912 = 1.5 * 2^23
}
// all threads of the CTA; the caller must __syncthreads() afterwards, then warp 0 calls tab_finish and syncs again
__device__ __forceinline__ Lookup tab_stage(uint8_t* sgen, uint32_t saddr, const uint32_t* __restrict__ table, int tid, int nthreads) {
  float* bp = reinterpret_cast<float*>(sgen);
  float4* piece = reinterpret_cast<float4*>(sgen + kCpbCells * 4);
  uint32_t* meta = reinterpret_cast<uint32_t*>(sgen + kCpbCells * 36);
  const float4* gcoef = reinterpret_cast<const float4*>(table + kTabCellCoef);
  const float* gbp = reinterpret_cast<const float*>(table + kTabCellBp);
  const uint16_t* gseg = reinterpret_cast<const uint16_t*>(table + kTabCellSeg);
  for (int i = tid; i < kCpbCells; i += nthreads) {
    const float4 e0 = __ldg(gcoef + i), e1 = __ldg(gcoef + kCpbCells + i);
    bp[i] = __ldg(gbp + i);
    piece[2 * i] = make_float4(e0.x, e0.y, e1.x, e1.y);
    piece[2 * i + 1] = make_float4(e0.z, e0.w, e1.z, e1.w);
    meta[i] = gseg[i];
  }
  Lookup L;
  L.bp = saddr;
  L.piece = saddr + kCpbCells * 4;
  L.meta = saddr + kCpbCells * 36;
  L.gtab = table;
  const float X = __uint_as_float(__ldg(table + 2)), inv = __uint_as_float(__ldg(table + 3));
  L.c1 = inv;
  L.c2 = X * inv;
  L.c2m = X * inv - 0.5f;
  L.bp_adj = L.bp - 0x4B400000u * 4u;
  L.piece_adj = L.piece - 0x4B400000u * 32u;
  return L;
}
__device__ __forceinline__ void tab_finish(uint8_t* sgen, int lane) {   // one warp: running count of flagged cells
  static_assert(kCpbCells % 32 == 0, "cells per lane");
  constexpr int kPer = kCpbCells / 32;                 // contiguous cells per lane
  const float* bp = reinterpret_cast<const float*>(sgen);
  uint32_t* meta = reinterpret_cast<uint32_t*>(sgen + kCpbCells * 36);
  uint32_t own = 0;
  for (int i = 0; i < kPer; ++i) {
    const float v = bp[lane * kPer + i];
    own += v != v ? 1u : 0u;
  }
  uint32_t incl = own;                                 // inclusive scan over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  uint32_t cnt = incl - own;                           // flagged cells before this lane's first cell
  for (int i = 0; i < kPer; ++i) {
    const int c = lane * kPer + i;
    const float v = bp[c];
    meta[c] = (meta[c] & 0xffffu) | (cnt << 16);
    cnt += v != v ? 1u : 0u;
  }
  if (lane == 31) meta[kCpbCells] = cnt << 16;
}
// largest |g| for which every |p| <= 1 + |g| stays inside the table domain
__device__ __forceinline__ float tab_gmax(const uint32_t* __restrict__ table) {
  return exp2f(__uint_as_float(__ldg(table + 2))) * 0.9995f - 2.0f;
}
static __device__ __noinline__ float4 lookup_slow(const uint32_t* gtab, int cell, float x, int* seg_out) {
  const uint16_t* cs = reinterpret_cast<const uint16_t*>(gtab + kTabCellSeg);
  const float* sbp = reinterpret_cast<const float*>(gtab + kTabSegBp);
  const float4* sc = reinterpret_cast<const float4*>(gtab + kTabSegCoef);
  int s = cs[cell];
  while (s < kCpbSegMax - 1 && x >= __ldg(sbp + s)) ++s;
  if (seg_out) *seg_out = s;
  return __ldg(sc + s);
}
// (a0, c0, a1, c1): bias_o (log2 domain) = a_o x + c_o.  x must lie inside the table domain (producers clamp g).
// kSeg: also return the global segment index of x.
template <bool kDirty, bool kSeg>
__device__ __forceinline__ float4 lookup2(const Lookup& L, float x, int& cell, int& seg) {
  const uint32_t raw = cell_raw(L, x);
  cell = (int)(raw - 0x4B400000u);
  const float bpv = lds_f32(L.bp_adj + raw * 4u);
  const bool hi = x >= bpv;
  uint32_t pa = L.piece_adj + raw * 32u;
  if (hi) pa += 16u;
  float4 e = lds_f32x4(pa);
  if (kSeg) seg = (lds_s32(L.meta + (uint32_t)cell * 4u) & 0xffff) + (hi ? 1 : 0);
  if (kDirty) {
    if (bpv != bpv) e = lookup_slow(L.gtab, cell, x, kSeg ? &seg : nullptr);
  }
  return e;
}
// ---- optional shared-memory copy of the first kSegSmem table segments (upper boundary, coefficients): the general
// lookup for x ranges that touch flagged cells or several segment boundaries, without going to global memory ----
constexpr int kSegSmem = 512;      // 10 KB; tables with more segments (hidden 32: at most 1088) take the callers' slow paths
constexpr uint32_t kSegSmemBytes = kSegSmem * 4 + kSegSmem * 16;
struct SegLookup {
  uint32_t bp, coef;     // shared-space addresses: float[kSegSmem], float4[kSegSmem]
  bool staged;           // false: the table has more segments than fit (callers fall back to lookup2<true, .>)
};
__device__ __forceinline__ SegLookup seg_stage(uint8_t* sgen, uint32_t saddr, const uint32_t* __restrict__ table, int tid, int nthreads,
                                               int limit = kSegSmem) {   // limit < kSegSmem: tests force the callers' slow paths
  SegLookup S;
  S.bp = saddr;
  S.coef = saddr + kSegSmem * 4;
  const int nseg = (int)__ldg(table);
  S.staged = nseg + 2 < min(limit, kSegSmem);  // callers read up to three boundaries / one coefficient pair past the last segment
  if (S.staged) {
    float* bp = reinterpret_cast<float*>(sgen);
    float4* coef = reinterpret_cast<float4*>(sgen + kSegSmem * 4);
    const float* gbp = reinterpret_cast<const float*>(table + kTabSegBp);
    const float4* gc = reinterpret_cast<const float4*>(table + kTabSegCoef);
    for (int i = tid; i < kSegSmem; i += nthreads) {
      bp[i] = i < nseg - 1 ? __ldg(gbp + i) : __int_as_float(0x7f800000);
      coef[i] = __ldg(gc + min(i, nseg - 1));
    }
  }
  return S;
}
__device__ __forceinline__ float4 lookup_seg(const Lookup& L, const SegLookup& S, float x, int& cell, int& seg) {
  cell = cell_index(L, x);
  int s = lds_s32(L.meta + (uint32_t)cell * 4u) & 0xffff;
  while (x >= lds_f32(S.bp + (uint32_t)s * 4u)) ++s;
  seg = s;
  return lds_f32x4(S.coef + (uint32_t)s * 16u);
}
// one head's (a, c) only: `hoff` = 8 * head selects the pair inside the 16-byte piece
template <bool kDirty, bool kSeg>
__device__ __forceinline__ float2 lookup2h(const Lookup& L, float x, uint32_t hoff, int& cell, int& seg) {
  const uint32_t raw = cell_raw(L, x);
  cell = (int)(raw - 0x4B400000u);
  const float bpv = lds_f32(L.bp_adj + raw * 4u);
  const bool hi = x >= bpv;
  uint32_t pa = L.piece_adj + raw * 32u + hoff;
  if (hi) pa += 16u;
  float2 e = lds_f32x2(pa);
  if (kSeg) seg = (lds_s32(L.meta + (uint32_t)cell * 4u) & 0xffff) + (hi ? 1 : 0);
  if (kDirty) {
    if (bpv != bpv) {
      const float4 f = lookup_slow(L.gtab, cell, x, kSeg ? &seg : nullptr);
      e = hoff ? make_float2(f.z, f.w) : make_float2(f.x, f.y);
    }
  }
  return e;
}
__device__ __forceinline__ int tab_dirty_between(const Lookup& L, int cell_lo, int cell_hi) {   // flagged cells in [lo, hi]
  return (lds_s32(L.meta + (uint32_t)(cell_hi + 1) * 4u) >> 16) - (lds_s32(L.meta + (uint32_t)cell_lo * 4u) >> 16);
}

// ---- host: tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// fp16 [B, rows, ld] tensor, box = 64 columns x box_rows rows, 128-byte swizzle, out-of-range rows read as zero
inline int make_map(CUtensorMap* m, const void* base, int B, int rows, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DML_EUNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DML_OK : DML_EINVAL;
}

}  // namespace tc
}  // namespace dml
