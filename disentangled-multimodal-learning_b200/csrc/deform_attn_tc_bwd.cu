// Backward of the fused deformable cross-attention on tcgen05 / TMEM / TMA (sm_100a).
//
// With P = softmax(S), S = scale q.k^T + bias(x) (see deform_attn_tc.cu), dO the incoming gradient (fp16, multiplied
// by the power-of-two loss scale s), D_i = sum_d dO_id O_id:
//     dP = dO V^T,  dS = P o (dP - D),  dQ = dS K,  dK = scale dS^T Q,  dV = P^T dO,
//     dg_j  = - sum_{h in group} sum_i dS_ij a_h(x_ij) / (|p_ij| + 1)              (bias slope a, p = seq_i - g_j)
//     segsum[s] = (sum dS_0, sum dS_0 x, sum dS_1, sum dS_1 x) over the pairs whose x lies in table segment s
// Without a workspace nothing of size n x n_kv is stored: both kernels recompute P from the saved log-sum-exp.  With the
// dS workspace (dml_deform_attn_bwd_ws_bytes) the dK/dV kernel also writes dS^T (fp16, [head][key][query]) and dQ = dS K
// becomes a plain streaming GEMM over it (deform_attn_dq_gemm_kernel) instead of a second recomputation.
//
//   deform_attn_dkv_tc_kernel  key-stationary (TMEM lane = key): one 128-key tile and both heads of the group per CTA,
//                              streaming 32-query tiles through a 4-stage TMA ring; S^T = K Q^T and dP^T = V dO^T are
//                              double-buffered in TMEM, P^T and dS^T (fp16) overwrite them in place and feed
//                              dV += P^T dO, dK += dS^T Q (TS MMAs, dO / Q tiles re-read as MN-major operands).  16
//                              elementwise warps: one head, one TMEM lane quarter and one 16-query half of the tile
//                              each.  A thread owns one key for all queries, so g_j is a register, dg_j accumulates
//                              privately and x_ij grows monotonically along the row: the thread carries its table
//                              segment from tile to tile (coefficients in registers, no per-position table access) and
//                              the per-segment sums leave as warp-reduced global reductions when the segment changes.
//                              CTAs take pieces of items from a cost-balanced work list (host wrapper below).
//   deform_attn_dq_gemm_kernel with the workspace: dQ = dS K streamed from the stored dS^T (MN-major A operand).
//   deform_attn_dq_tc_kernel   without it: query-stationary (TMEM lane = query), two 128-query groups per CTA alternate
//                              on the tensor pipe; per 32-key tile S = Q K^T and dP = dO V^T (SS MMAs), dS (fp16)
//                              overwrites S in TMEM, dQ += dS K (TS MMA, K as an MN-major operand).
#include <math.h>
#include <stdlib.h>

#include "../../include/dml_b200.h"
#include "tc_common.cuh"

namespace dml {
namespace tc {

constexpr int kD = 64;
constexpr float kLseOff = 1.0e30f;   // log-sum-exp stand-in for rows past the end: P = exp2(. - 1e30) = 0

struct BwdParams {
  const float* g;        // [(B G), n_kv]
  const uint32_t* table;
  const float* lse;      // [B, H, n] (log2 domain)
  const float* dsum;     // [B, H, n]  D (scaled by s like dO)
  const float* dscale;   // device float[2] = (s, 1/s)
  float* dq;             // fp32 [B, n, H*64]     (unscaled by `scale`, as dml_deform_attn_bwd)
  float* dk; float* dv;  // fp32 [B, n_kv, H*64]
  float* dg;             // [(B G), n_kv]  accumulated (zeroed by the host wrapper)
  float* segsum;         // [kCpbSegMax][4] accumulated (zeroed by the host wrapper)
  int B, H, n, n_kv, n_seq;
  float scale;
  h16* ds_ws;            // optional fp16 [(B H), n_kv_pad, n_pad]: dS^T (times s) written by the dK/dV kernel
  int n_pad, n_kv_pad;
  int seg_limit;         // dK/dV kernel: tables with >= seg_limit - 2 segments are not staged (debug knob, default kSegSmem)
  int qsplit;            // dK/dV kernel without a work list: the query tiles of one item are shared by this many CTAs
  long long* trace;      // debug: per-tile clock64() stamps of CTA (0,0,0) of the dQ kernel (nullptr = off)
};
// Work list of the dK/dV kernel (kernel parameter, read through the constant bank): CTA i takes query tiles [t0, t1) of
// item `item` = key block + nkb * (head pair + G * batch).  n == 0: no list, blockIdx.x = item * qsplit + part.
struct DkvWork { uint16_t item, t0, t1, pad; };
constexpr int kDkvWorkMax = 320;
struct DkvWorkList {
  int n;
  DkvWork e[kDkvWorkMax];
};
#ifdef DML_TEST_KNOBS      // test-only build (libdml_b200_test.so): the product library has no mutable global state
static long long* g_trace = nullptr;
static int g_seg_limit = kSegSmem;
#endif

// D[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]   (O fp32 as written by the forward, dO the fp16 tensor the MMAs consume)
__global__ void __launch_bounds__(256) bwd_prep_kernel(const float* __restrict__ o, const h16* __restrict__ d_o, int B, int n,
                                                       int H, int ld, float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int per_lane = (H * kD) / 32;  // 16 for H = 8
  const int lanes_per_head = kD / per_lane;
  for (int row = warp; row < B * n; row += nwarps) {
    const float* po = o + (size_t)row * ld + lane * per_lane;
    const h16* pd = d_o + (size_t)row * ld + lane * per_lane;
    float s = 0.f;
    for (int e = 0; e < per_lane; e += 8) {
      const float4 a0 = *reinterpret_cast<const float4*>(po + e), a1 = *reinterpret_cast<const float4*>(po + e + 4);
      uint4 c = *reinterpret_cast<const uint4*>(pd + e);
      const __half2* c2 = reinterpret_cast<const __half2*>(&c);
      s = fmaf(a0.x, __low2float(c2[0]), s); s = fmaf(a0.y, __high2float(c2[0]), s);
      s = fmaf(a0.z, __low2float(c2[1]), s); s = fmaf(a0.w, __high2float(c2[1]), s);
      s = fmaf(a1.x, __low2float(c2[2]), s); s = fmaf(a1.y, __high2float(c2[2]), s);
      s = fmaf(a1.z, __low2float(c2[3]), s); s = fmaf(a1.w, __high2float(c2[3]), s);
    }
    for (int off = lanes_per_head >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((lane % lanes_per_head) == 0) {
      const int hh = lane / lanes_per_head, bb = row / n, i = row % n;
      dsum[((size_t)bb * H + hh) * n + i] = s;
    }
  }
}

// =================================================================================================================
// dQ kernel
// =================================================================================================================
namespace dqk {
constexpr int kBM = 128, kGroups = 2, kBN = 32, kStages = 3;
constexpr int kEwWarps = 16;                      // 8 per group: 4 lane quarters x 2 key halves of each 32-key tile
constexpr int kThreads = 32 * (kEwWarps + 3);    // + TMA producer + one MMA-issuing warp per group
constexpr uint32_t kTileQ = kBM * kD * 2;       // 16 KB
constexpr uint32_t kTileKV = kBN * kD * 2;      // 4 KB
constexpr uint32_t kStageBytes = 4 * kTileKV;   // K0 K1 V0 V1
constexpr uint32_t kOffQ = 0;                                           // [group][head]
constexpr uint32_t kOffDO = kOffQ + kGroups * 2 * kTileQ;               // [group][head]
constexpr uint32_t kOffKV = kOffDO + kGroups * 2 * kTileQ;              // [stage]{K0,K1,V0,V1}
constexpr uint32_t kOffG = kOffKV + kStages * kStageBytes;              // [stage][32 g, gmin, gmax, pad]
constexpr uint32_t kGStride = 40 * 4;
constexpr uint32_t kOffTab = kOffG + kStages * kGStride + 32;           // 16-B aligned
constexpr uint32_t kOffBar = kOffTab + kTabSmemBytes;
constexpr int kBarQ = 0, kBarKvFull = 1, kBarKvEmpty = kBarKvFull + kStages, kBarSFull = kBarKvEmpty + kStages,
              kBarPFull = kBarSFull + kGroups, kNumBars = kBarPFull + kGroups;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemPtr + 16 + 1024;
static_assert(kOffTab % 16 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr uint32_t kIdescSD = idesc_f16(128, kBN, false, false);   // S = Q K^T, dP = dO V^T
constexpr uint32_t kIdescDQ = idesc_f16(128, 64, false, true);     // dQ += dS K   (K MN-major)
}  // namespace dqk

// 16 keys (one half of a 32-key tile) of one query row, both heads: dS = P (dP - D) as fp16 pairs written over the
// first 8 of the 16 S columns the warp owns.  tS = TMEM address of the warp's first S column, gsa = its first g.
template <bool kMasked, bool kDirty>
__device__ __forceinline__ void dq_sweep(const Lookup& L, uint32_t tS, uint32_t gsa, float s_i, float sc2, float lse0,
                                         float lse1, float d0, float d1, int jrem) {
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t a[8], bq[8], pa[8], pb[8];
    tmem_ld8(tS + c * 8, a);
    tmem_ld8(tS + 32 + c * 8, bq);
    tmem_ld8(tS + 64 + c * 8, pa);
    tmem_ld8(tS + 96 + c * 8, pb);
    float gq[8];
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 t = lds_f32x4(gsa + (uint32_t)(c * 8 + e) * 4);
      gq[e] = t.x; gq[e + 1] = t.y; gq[e + 2] = t.z; gq[e + 3] = t.w;
    }
    tmem_ld_fence();
    reg_fence(a); reg_fence(bq); reg_fence(pa); reg_fence(pb);
    uint32_t w0[4], w1[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float v0[2], v1[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float x = cpb_x(s_i - gq[e + u]);
        int cdummy, sdummy;
        const float4 t = lookup2<kDirty, false>(L, x, cdummy, sdummy);
        const float p0 = ex2(fmaf(__uint_as_float(a[e + u]), sc2, fmaf(t.x, x, t.y)) - lse0);
        const float p1 = ex2(fmaf(__uint_as_float(bq[e + u]), sc2, fmaf(t.z, x, t.w)) - lse1);
        v0[u] = p0 * (__uint_as_float(pa[e + u]) - d0);
        v1[u] = p1 * (__uint_as_float(pb[e + u]) - d1);
        if (kMasked && c * 8 + e + u >= jrem) { v0[u] = 0.f; v1[u] = 0.f; }
      }
      w0[e >> 1] = pack_f16(v0[0], v0[1]);
      w1[e >> 1] = pack_f16(v1[0], v1[1]);
    }
    tmem_st4(tS + c * 4, w0);
    tmem_st4(tS + 32 + c * 4, w1);
  }
}

__global__ void __launch_bounds__(dqk::kThreads, 1)
deform_attn_dq_tc_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mdo,
                         const __grid_constant__ CUtensorMap mk, const __grid_constant__ CUtensorMap mv,
                         const BwdParams p) {
  using namespace dqk;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int i0 = blockIdx.x * (kGroups * kBM), grp = blockIdx.y, b = blockIdx.z;
  const int G = p.H / 2;
  const int ntiles = cdiv(p.n_kv, kBN);
  auto bar = [&](int i) { return sbase + kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kOffTmemPtr);

  if (tid == 0) {
    mbar_init(bar(kBarQ), 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarKvFull + s), 32); mbar_init(bar(kBarKvEmpty + s), kGroups); }
    for (int g = 0; g < kGroups; ++g) { mbar_init(bar(kBarSFull + g), 1); mbar_init(bar(kBarPFull + g), 2 * kBM); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEwWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const Lookup L = tab_stage(sgen + kOffTab, sbase + kOffTab, p.table, tid, kThreads);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tab_finish(sgen + kOffTab, lane);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kEwWarps) {
    // ---- TMA producer ----
    if (lane == 0) {
      mbar_expect_tx(bar(kBarQ), 2 * kGroups * 2 * kTileQ);
      for (int g = 0; g < kGroups; ++g)
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(sbase + kOffQ + (g * 2 + h) * kTileQ, &mq, bar(kBarQ), (grp * 2 + h) * kD, i0 + g * kBM, b);
          tma_load_3d(sbase + kOffDO + (g * 2 + h) * kTileQ, &mdo, bar(kBarQ), (grp * 2 + h) * kD, i0 + g * kBM, b);
        }
    }
    const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
    const float gb_max = tab_gmax(p.table);
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      mbar_wait(bar(kBarKvEmpty + st), ((j / kStages) & 1) ^ 1);
      const uint32_t dst = sbase + kOffKV + st * kStageBytes;
      if (lane == 0) {
        mbar_expect_tx_only(bar(kBarKvFull + st), kStageBytes);
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(dst + h * kTileKV, &mk, bar(kBarKvFull + st), (grp * 2 + h) * kD, j * kBN, b);
          tma_load_3d(dst + (2 + h) * kTileKV, &mv, bar(kBarKvFull + st), (grp * 2 + h) * kD, j * kBN, b);
        }
      }
      float* gs = reinterpret_cast<float*>(sgen + kOffG + st * kGStride);
      const float g0 = fminf(fmaxf(__ldg(gb + min(j * kBN + lane, p.n_kv - 1)), -gb_max), gb_max);
      gs[lane] = g0;
      const float gmn = -warp_max(-g0), gmx = warp_max(g0);
      if (lane == 0) { gs[32] = gmn; gs[33] = gmx; }
      mbar_arrive(bar(kBarKvFull + st));
    }
  } else if (warp > kEwWarps) {
    // ---- MMA issuers: warp kEwWarps+1 -> group 0, kEwWarps+2 -> group 1 (all lanes run the loop, one elected lane issues) ----
    const int g = warp - (kEwWarps + 1);
    const bool leader = elect_one();
    mbar_wait(bar(kBarQ), 0);
    auto issue_sd = [&](int j) {
      const int st = j % kStages;
      const uint32_t kv = sbase + kOffKV + st * kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t dq_ = smem_desc(sbase + kOffQ + (g * 2 + h) * kTileQ), dk_ = smem_desc(kv + h * kTileKV);
        const uint64_t dd_ = smem_desc(sbase + kOffDO + (g * 2 + h) * kTileQ), dv_ = smem_desc(kv + (2 + h) * kTileKV);
        const uint32_t ds = tmem + g * 256 + h * 32, dp = tmem + g * 256 + 64 + h * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(ds, dq_ + 2 * k, dk_ + 2 * k, kIdescSD, k > 0, leader);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(dp, dd_ + 2 * k, dv_ + 2 * k, kIdescSD, k > 0, leader);
      }
      tc_commit(bar(kBarSFull + g), leader);
    };
    mbar_wait(bar(kBarKvFull + 0), 0);
    tc_fence_after();
    issue_sd(0);
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      const uint32_t kv = sbase + kOffKV + st * kStageBytes;
      mbar_wait(bar(kBarPFull + g), j & 1);
      tc_fence_after();
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
      if (tr) p.trace[(j * 2 + g) * 4 + 2] = clock64();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t db = smem_desc(kv + h * kTileKV);                       // K tile as [key][d]: MN-major B
        const uint32_t d = tmem + g * 256 + 128 + h * 64, a = tmem + g * 256 + h * 32;
#pragma unroll
        for (int k = 0; k < 2; ++k) mma_ts(d, a + 16 * k, db + 128 * k, kIdescDQ, (j > 0) || (k > 0), leader);   // dS of keys 16k.. sits in columns 16k..16k+7
      }
      tc_commit(bar(kBarKvEmpty + st), leader);
      if (j + 1 < ntiles) {
        mbar_wait(bar(kBarKvFull + (j + 1) % kStages), ((j + 1) / kStages) & 1);
        tc_fence_after();
        issue_sd(j + 1);
      } else {
        tc_commit(bar(kBarSFull + g), leader);
      }
      if (tr) p.trace[(j * 2 + g) * 4 + 3] = clock64();
    }
  } else {
    // ---- elementwise warps: group = warp / 8, TMEM lane quarter = warp & 3, key half of every tile = (warp >> 2) & 1 ----
    const int g = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int row = (warp & 3) * 32 + lane;
    const int gi = i0 + g * kBM + row;
    const bool rv = gi < p.n;
    const uint32_t tbase = tmem + g * 256 + (((uint32_t)(warp & 3) * 32u) << 16);
    const int h0 = grp * 2;
    const size_t ro = ((size_t)b * p.H + h0) * p.n + min(gi, p.n - 1);
    const float lse0 = rv ? __ldg(p.lse + ro) : kLseOff, lse1 = rv ? __ldg(p.lse + ro + p.n) : kLseOff;
    const float d0 = rv ? __ldg(p.dsum + ro) : 0.f, d1 = rv ? __ldg(p.dsum + ro + p.n) : 0.f;
    const float s_i = seq_pos(min(gi, p.n_seq - 1), p.n_seq);
    const float sc2 = p.scale * kLog2e;

    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      const uint32_t gsa = sbase + kOffG + st * kGStride;
      mbar_wait(bar(kBarKvFull + st), (j / kStages) & 1);
      mbar_wait(bar(kBarSFull + g), j & 1);
      tc_fence_after();
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (warp & 7) == 0 && lane == 0;
      if (tr) p.trace[(j * 2 + g) * 4 + 0] = clock64();
      const int jrem = p.n_kv - j * kBN - half * 16;            // valid keys in this warp's half of the tile (may be <= 0)
      int ndirty;
      {
        const float xlo = cpb_x(s_i - lds_f32(gsa + 33 * 4)), xhi = cpb_x(s_i - lds_f32(gsa + 32 * 4));
        const int clo = cell_index(L, xlo), chi = cell_index(L, xhi);
        ndirty = tab_dirty_between(L, clo, chi);
      }
      const bool dirty = __any_sync(0xffffffffu, ndirty != 0);
      const uint32_t tS = tbase + half * 16, gh = gsa + half * 64;
      if (jrem >= 16) {
        if (!dirty) dq_sweep<false, false>(L, tS, gh, s_i, sc2, lse0, lse1, d0, d1, jrem);
        else dq_sweep<false, true>(L, tS, gh, s_i, sc2, lse0, lse1, d0, d1, jrem);
      } else {
        dq_sweep<true, true>(L, tS, gh, s_i, sc2, lse0, lse1, d0, d1, jrem);
      }
      if (tr) p.trace[(j * 2 + g) * 4 + 1] = clock64();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(kBarPFull + g));
    }

    // ---- epilogue: dQ / s -> global (this warp: 32 of the 64 columns of each head) ----
    mbar_wait(bar(kBarSFull + g), ntiles & 1);
    tc_fence_after();
    const float inv_s = __ldg(p.dscale + 1);
    float* ob = p.dq + ((size_t)b * p.n + gi) * (p.H * kD) + h0 * kD;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = (c >> 1) * 64 + half * 32 + (c & 1) * 16;     // head c>>1, columns half*32 + (c&1)*16 ..
      uint32_t a[16];
      tmem_ld16(tbase + 128 + col, a);
      tmem_ld_wait(a);
      if (rv) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(ob + col + e) =
              make_float4(__uint_as_float(a[e]) * inv_s, __uint_as_float(a[e + 1]) * inv_s,
                          __uint_as_float(a[e + 2]) * inv_s, __uint_as_float(a[e + 3]) * inv_s);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEwWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// =================================================================================================================
// dK / dV / dg / segment-sum kernel
// =================================================================================================================
namespace dkvk {
constexpr int kBK = 128, kBI = 32, kStages = 4;
constexpr int kEwWarps = 16;                      // per head 8: 4 TMEM lane quarters x 2 query halves of each 32-query tile
constexpr int kThreads = 32 * (kEwWarps + 2);    // + TMA producer + MMA issuer
constexpr uint32_t kTileKV = kBK * kD * 2;      // 16 KB
constexpr uint32_t kTileQ = kBI * kD * 2;       // 4 KB
constexpr uint32_t kStageBytes = 4 * kTileQ;    // Q0 Q1 dO0 dO1
constexpr uint32_t kOffKV = 0;                                          // K0 K1 V0 V1 (resident)
constexpr uint32_t kOffIn = kOffKV + 4 * kTileKV;                       // [stage]{Q0,Q1,dO0,dO1}
constexpr uint32_t kOffRow = kOffIn + kStages * kStageBytes;            // [stage]{seq[32], lse0[32], lse1[32], D0[32], D1[32]}
constexpr uint32_t kRowStride = 5 * 32 * 4;
constexpr uint32_t kOffTab = kOffRow + kStages * kRowStride;
constexpr uint32_t kOffSeg = kOffTab + kTabSmemBytes;                   // shared-memory segment arrays (seg_stage)
constexpr uint32_t kOffBar = kOffSeg + kSegSmemBytes;
constexpr int kBarKv = 0, kBarInFull = 1, kBarInEmpty = kBarInFull + kStages, kBarSFull = kBarInEmpty + kStages,
              kBarPFull = kBarSFull + 2, kBarAcc = kBarPFull + 2, kNumBars = kBarAcc + 1;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemPtr + 16 + 1024;
static_assert(kOffTab % 16 == 0 && kOffSeg % 16 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr uint32_t kIdescSD = idesc_f16(128, kBI, false, false);   // S^T = K Q^T, dP^T = V dO^T
constexpr uint32_t kIdescAcc = idesc_f16(128, 64, false, true);    // dV += P^T dO, dK += dS^T Q  (dO / Q MN-major)
}  // namespace dkvk

struct SegRun {      // run-length state of one thread: sums of the current segment for the thread's head
  int seg;
  float a, b;
};
struct SegSink {     // where finished segment sums go: the global accumulator (fire-and-forget reductions at L2), un-scaled
  float* sum;
  float inv_s;
};
__device__ __forceinline__ void seg_flush(const SegSink& ssum, SegRun& r, int head) {
  if (r.seg >= 0 && (r.a != 0.f || r.b != 0.f)) {
    atomicAdd(ssum.sum + 4 * r.seg + 2 * head, r.a * ssum.inv_s);
    atomicAdd(ssum.sum + 4 * r.seg + 2 * head + 1, r.b * ssum.inv_s);
  }
  r.a = r.b = 0.f;
}

// Warp-cooperative flush (all 32 lanes call it): lanes with `need` hand (seg, a, b) over; lanes naming the same segment -
// neighbouring keys almost always do - are summed with shuffles and one lane issues the two reductions.
__device__ __forceinline__ void seg_flush_warp(const SegSink& ssum, bool need, int seg, float a, float b, int head, int lane) {
  need = need && seg >= 0 && (a != 0.f || b != 0.f);
  uint32_t pend = __ballot_sync(0xffffffffu, need);
  while (pend) {
    const int s = __shfl_sync(0xffffffffu, seg, __ffs(pend) - 1);
    const bool mine = need && seg == s;
    float va = mine ? a : 0.f, vb = mine ? b : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      va += __shfl_xor_sync(0xffffffffu, va, o);
      vb += __shfl_xor_sync(0xffffffffu, vb, o);
    }
    if (lane == 0) {
      atomicAdd(ssum.sum + 4 * s + 2 * head, va * ssum.inv_s);
      atomicAdd(ssum.sum + 4 * s + 2 * head + 1, vb * ssum.inv_s);
    }
    pend &= ~__ballot_sync(0xffffffffu, mine);
  }
}

// 16 queries of one key row, one head.  S^T / dP^T (fp32, TMEM) -> P^T / dS^T (fp16 pairs, in place); accumulates
// dg and the per-segment sums.  x grows along the row, so the segment index only ever increases.
// The thread therefore knows its segment: bnd = (xb, xb2, xb3) are the next three boundaries after seg_first.
//   kMode 0  the 16 positions lie in at most two adjacent segments [seg_first, seg_first + 1] split at xb: the two
//            segments' coefficients (clo, chi) sit in registers, a position costs one compare and two selects - no table
//            access at all - and the sums go to two register buckets (the common case);
//   kMode 3  no lane of the warp meets a boundary in its 16 positions (most tiles away from the diagonal): one pair of
//            coefficients, no compare / select, one bucket, and the slope leaves the dg sum;
//   kMode 1  some lane of the warp crosses two or three boundaries: segment = seg_first + number of boundaries passed,
//            coefficients from the shared-memory segment array, four register buckets;
//   kMode 2  anything else (table with more than kSegSmem segments, > 3 boundaries in 16 positions): per-position
//            lookup through the cell index / global table, run-length merged with per-thread reductions.
// tS = TMEM address of the warp's first S^T column of this head (dP^T sits 64 columns further).
template <bool kKeyMasked, int kMode>
__device__ __forceinline__ void dkv_sweep(const Lookup& L, const SegLookup& SL, uint32_t tS, uint32_t rowa, int head, float g_j,
                                          bool key_valid, float sc2, int seg_first, int seg_last, float3 bnd, float2 clo, float2 chi,
                                          float& dgacc, SegRun& run, const SegSink& ssum, h16* dsp, long long* trc = nullptr) {
  const int lane = threadIdx.x & 31;
  const uint32_t hoff = head ? 8u : 0u;
  const float xb = bnd.x;
  float ba[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};   // kMode 0: [0] whole tile, [1] at or above xb; kMode 1: per segment
  float dgflat = 0.f;                                                  // kMode 3: sum of dS / (|p| + 1)
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {                           // two sub-chunks of 8 queries
    uint32_t a[8], pa[8];
    tmem_ld8(tS + c * 8, a);
    tmem_ld8(tS + 64 + c * 8, pa);
    float sq[8], l0[8], d0[8];
    const uint32_t ra = rowa + c * 32;
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      float4 t = lds_f32x4(ra + e * 4);
      sq[e] = t.x; sq[e + 1] = t.y; sq[e + 2] = t.z; sq[e + 3] = t.w;
      t = lds_f32x4(ra + 128 + head * 128 + e * 4);
      l0[e] = t.x; l0[e + 1] = t.y; l0[e + 2] = t.z; l0[e + 3] = t.w;
      t = lds_f32x4(ra + 384 + head * 128 + e * 4);
      d0[e] = t.x; d0[e + 1] = t.y; d0[e + 2] = t.z; d0[e + 3] = t.w;
    }
    tmem_ld_fence();
    reg_fence(a); reg_fence(pa);
    uint32_t wp[4], ws[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float pp[2], dd[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float pr = sq[e + u] - g_j;
        const float qa = fabsf(pr) + 1.0f;
        const float x = copysignf(__log2f(qa), pr);
        int cell, seg = 0;
        const bool above = x >= xb;
        float2 t;                                          // this head's (slope, intercept)
        if (kMode == 3) {
          t = clo;
        } else if (kMode == 0) {
          t = above ? chi : clo;
        } else if (kMode == 1) {
          seg = (above ? 1 : 0) + (x >= bnd.y ? 1 : 0) + (x >= bnd.z ? 1 : 0);      // relative to seg_first
          t = lds_f32x2(SL.coef + (uint32_t)(seg_first + seg) * 16u + hoff);
        } else {
          t = lookup2h<true, true>(L, x, hoff, cell, seg);
        }
        const float slope = t.x, icpt = t.y;
        float p0 = ex2(fmaf(__uint_as_float(a[e + u]), sc2, fmaf(slope, x, icpt)) - l0[e + u]);
        float s0 = p0 * (__uint_as_float(pa[e + u]) - d0[e + u]);
        if (kKeyMasked && !key_valid) { p0 = 0.f; s0 = 0.f; }
        pp[u] = p0; dd[u] = s0;
        // d bias / d g_j = -a / (|p| + 1)
        if (kMode == 3) {                                   // one slope for the whole sweep: it multiplies the sum afterwards
          dgflat = fmaf(s0, rcp_approx(qa), dgflat);
          ba[0] += s0; bb[0] = fmaf(s0, x, bb[0]);
          continue;
        }
        dgacc = fmaf(s0 * slope, -rcp_approx(qa), dgacc);
        if (kMode == 2) {
          if (seg != run.seg) { seg_flush(ssum, run, head); run.seg = seg; }
          run.a += s0; run.b = fmaf(s0, x, run.b);
        } else if (kMode == 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float m = seg == q ? s0 : 0.f;    // select, not a branch: the positions stay interleaved
            ba[q] += m; bb[q] = fmaf(m, x, bb[q]);
          }
        } else {
          ba[0] += s0; bb[0] = fmaf(s0, x, bb[0]);
          if (above) { ba[1] += s0; bb[1] = fmaf(s0, x, bb[1]); }
        }
      }
      wp[e >> 1] = pack_f16(pp[0], pp[1]);
      ws[e >> 1] = pack_f16(dd[0], dd[1]);
    }
    tmem_st4(tS + c * 4, wp);
    tmem_st4(tS + 64 + c * 4, ws);
    if (dsp) *reinterpret_cast<uint4*>(dsp + c * 8) = make_uint4(ws[0], ws[1], ws[2], ws[3]);   // dS^T row of this key, 8 queries
  }
  if (kMode == 0) { ba[0] -= ba[1]; bb[0] -= bb[1]; }      // [0] below the boundary, [1] at or above it
  if (kMode == 3) dgacc = fmaf(-clo.x, dgflat, dgacc);
  if (trc) *trc = clock64();
  if (kMode != 2) {
    // hand the buckets to the running sums.  Completed segments of this lane, in order: the old run when it is not
    // seg_first (else it merges into bucket 0), then buckets 0 .. span-1; bucket `span` becomes the new run.  Round j
    // flushes every lane's j-th completed segment warp-cooperatively, so a tile costs as many rounds as the worst lane
    // has completed segments (one in the usual boundary-crossing tile).
    constexpr int kB = kMode == 1 ? 4 : kMode == 3 ? 1 : 2;
    const int span = seg_last - seg_first;
    const bool old = seg_first != run.seg;
    if (!old) { ba[0] += run.a; bb[0] += run.b; }
    const int mine = (old ? 1 : 0) + span;
    if (__any_sync(0xffffffffu, mine > 0)) {
      const int rounds = __reduce_max_sync(0xffffffffu, mine);
      for (int j = 0; j < rounds; ++j) {
        const int idx = j - (old ? 1 : 0);               // -1: the old run, else bucket idx
        float fa = run.a, fb = run.b;
#pragma unroll
        for (int q = 0; q < kB - 1; ++q)
          if (idx == q) { fa = ba[q]; fb = bb[q]; }
        seg_flush_warp(ssum, j < mine, idx < 0 ? run.seg : seg_first + idx, fa, fb, head, lane);
      }
    }
    run.seg = seg_last;
    run.a = ba[0]; run.b = bb[0];
#pragma unroll
    for (int q = 1; q < kB; ++q)
      if (span == q) { run.a = ba[q]; run.b = bb[q]; }
  }
}

__global__ void __launch_bounds__(dkvk::kThreads, 1)
deform_attn_dkv_tc_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mdo,
                          const __grid_constant__ CUtensorMap mk, const __grid_constant__ CUtensorMap mv,
                          const BwdParams p, const __grid_constant__ DkvWorkList wl) {
  using namespace dkvk;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  // A CTA takes the query tiles [t_begin, ntiles) of one item (batch, head pair, 128-key block).  With 128 items on 148
  // SMs a one-CTA-per-item launch leaves 20 SMs idle, so the host cuts the items' tile ranges into pieces of about equal
  // estimated cost, lists them longest first (the work list) and the pieces of an item add their dK / dV into the zeroed
  // outputs with float4 reductions; without a list blockIdx.x = item * qsplit + part.
  int item, t_begin, ntiles;      // ntiles = END of this CTA's tile range
  if (wl.n > 0) {
    const DkvWork w = wl.e[blockIdx.x];
    item = w.item; t_begin = w.t0; ntiles = w.t1;
  } else {
    const int part = blockIdx.x % p.qsplit;
    item = blockIdx.x / p.qsplit;
    t_begin = (int)((long long)cdiv(p.n, kBI) * part / p.qsplit);
    ntiles = (int)((long long)cdiv(p.n, kBI) * (part + 1) / p.qsplit);
  }
  const bool whole = t_begin == 0 && ntiles == cdiv(p.n, kBI);      // the only CTA of its item: plain stores
  const int G = p.H / 2, nkb = cdiv(p.n_kv, kBK);
  const int j0 = (item % nkb) * kBK, grp = (item / nkb) % G, b = item / (nkb * G);
  const int h0 = grp * 2;
  auto bar = [&](int i) { return sbase + kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kOffTmemPtr);

  if (tid == 0) {
    mbar_init(bar(kBarKv), 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarInFull + s), 32); mbar_init(bar(kBarInEmpty + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar(kBarSFull + s), 1); mbar_init(bar(kBarPFull + s), 32 * kEwWarps); }
    mbar_init(bar(kBarAcc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEwWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const Lookup L = tab_stage(sgen + kOffTab, sbase + kOffTab, p.table, tid, kThreads);
  const SegLookup SL = seg_stage(sgen + kOffSeg, sbase + kOffSeg, p.table, tid, kThreads, p.seg_limit);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tab_finish(sgen + kOffTab, lane);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kEwWarps) {
    // ---- TMA producer ----
    if (lane == 0) {
      mbar_expect_tx(bar(kBarKv), 4 * kTileKV);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(sbase + kOffKV + h * kTileKV, &mk, bar(kBarKv), (h0 + h) * kD, j0, b);
        tma_load_3d(sbase + kOffKV + (2 + h) * kTileKV, &mv, bar(kBarKv), (h0 + h) * kD, j0, b);
      }
    }
    const float* lb = p.lse + ((size_t)b * p.H + h0) * p.n;
    const float* db = p.dsum + ((size_t)b * p.H + h0) * p.n;
    for (int t = t_begin; t < ntiles; ++t) {
      const int it = t - t_begin, st = it % kStages;
      mbar_wait_relaxed(bar(kBarInEmpty + st), ((it / kStages) & 1) ^ 1);
      const uint32_t dst = sbase + kOffIn + st * kStageBytes;
      if (lane == 0) {
        mbar_expect_tx_only(bar(kBarInFull + st), kStageBytes);
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(dst + h * kTileQ, &mq, bar(kBarInFull + st), (h0 + h) * kD, t * kBI, b);
          tma_load_3d(dst + (2 + h) * kTileQ, &mdo, bar(kBarInFull + st), (h0 + h) * kD, t * kBI, b);
        }
      }
      float* rs = reinterpret_cast<float*>(sgen + kOffRow + st * kRowStride);
      const int i = t * kBI + lane;
      const bool iv = i < p.n;
      const int ic = min(i, p.n - 1);
      rs[lane] = seq_pos(min(i, p.n_seq - 1), p.n_seq);
      rs[32 + lane] = iv ? __ldg(lb + ic) : kLseOff;
      rs[64 + lane] = iv ? __ldg(lb + p.n + ic) : kLseOff;
      rs[96 + lane] = iv ? __ldg(db + ic) : 0.f;
      rs[128 + lane] = iv ? __ldg(db + p.n + ic) : 0.f;
      mbar_arrive(bar(kBarInFull + st));
    }
  } else if (warp == kEwWarps + 1) {
    // ---- MMA issuer (all lanes run the loop under uniform control flow, one elected lane issues) ----
    {
      const bool leader = elect_one();
      mbar_wait(bar(kBarKv), 0);
      auto issue_sd = [&](int it) {
        const int st = it % kStages, buf = it & 1;
        const uint32_t in = sbase + kOffIn + st * kStageBytes;
        for (int h = 0; h < 2; ++h) {
          const uint64_t dk_ = smem_desc(sbase + kOffKV + h * kTileKV), dq_ = smem_desc(in + h * kTileQ);
          const uint64_t dv_ = smem_desc(sbase + kOffKV + (2 + h) * kTileKV), dd_ = smem_desc(in + (2 + h) * kTileQ);
          const uint32_t ds = tmem + buf * 128 + h * 32, dp = tmem + buf * 128 + 64 + h * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(ds, dk_ + 2 * k, dq_ + 2 * k, kIdescSD, k > 0, leader);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(dp, dv_ + 2 * k, dd_ + 2 * k, kIdescSD, k > 0, leader);
        }
        tc_commit(bar(kBarSFull + buf), leader);
      };
      mbar_wait(bar(kBarInFull + 0), 0);
      tc_fence_after();
      issue_sd(0);
#ifdef DML_TRACE
      const bool tr = p.trace && blockIdx.x == 0 && lane == 0;
#else
      constexpr bool tr = false;
#endif
      const int nit = ntiles - t_begin;
      for (int t = 0; t < nit; ++t) {                    // t counts this CTA's tiles from here on
        const int st = t % kStages, buf = t & 1;
        if (t + 1 < nit) {
          mbar_wait(bar(kBarInFull + (t + 1) % kStages), ((t + 1) / kStages) & 1);
          tc_fence_after();
          issue_sd(t + 1);
        }
        if (tr) p.trace[t * 8 + 7] = clock64();
        mbar_wait_relaxed(bar(kBarPFull + buf), (t >> 1) & 1);
        tc_fence_after();
        if (tr) p.trace[t * 8 + 5] = clock64();
        const uint32_t in = sbase + kOffIn + st * kStageBytes;
        for (int h = 0; h < 2; ++h) {
          const uint64_t dq_ = smem_desc(in + h * kTileQ), dd_ = smem_desc(in + (2 + h) * kTileQ);   // [query][d]: MN-major B
          const uint32_t pT = tmem + buf * 128 + h * 32, sT = tmem + buf * 128 + 64 + h * 32;
          const uint32_t dV = tmem + 256 + h * 64, dK = tmem + 384 + h * 64;
#pragma unroll
          for (int k = 0; k < 2; ++k) mma_ts(dV, pT + 16 * k, dd_ + 128 * k, kIdescAcc, (t > 0) || (k > 0), leader);
#pragma unroll
          for (int k = 0; k < 2; ++k) mma_ts(dK, sT + 16 * k, dq_ + 128 * k, kIdescAcc, (t > 0) || (k > 0), leader);
        }
        tc_commit(bar(kBarInEmpty + st), leader);
        if (tr) p.trace[t * 8 + 6] = clock64();
      }
      tc_commit(bar(kBarAcc), leader);
    }
  } else {
    // ---- elementwise warps: warp w -> head w >> 3 of the pair, TMEM lanes 32 (w & 3).., queries 16 ((w >> 2) & 1).. of
    //      each 32-query tile (the two heads' warps repeat the bias lookup but need no synchronisation) ----
    const int head = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int row = (warp & 3) * 32 + lane;
    const int gj = j0 + row;
    const bool kvld = gj < p.n_kv;
    const bool key_masked = (j0 + kBK) > p.n_kv;       // CTA-uniform
    const float gmaxv = tab_gmax(p.table);
    const float g_j = fminf(fmaxf(__ldg(p.g + (size_t)(b * G + grp) * p.n_kv + min(gj, p.n_kv - 1)), -gmaxv), gmaxv);
    const uint32_t lane_off = ((uint32_t)(warp & 3) * 32u) << 16;
    const float sc2 = p.scale * kLog2e;
    float dgacc = 0.f;
    h16* const ds_row =
        p.ds_ws ? p.ds_ws + ((size_t)(b * p.H + h0 + head) * p.n_kv_pad + gj) * p.n_pad + t_begin * kBI + half * 16 : nullptr;
    SegRun run;
    run.seg = -1;
    run.a = run.b = 0.f;
    const float inv_s = __ldg(p.dscale + 1);
    const SegSink ssum{p.segsum, inv_s};

#ifdef DML_TRACE      // clock64() stamps of CTA 0 for scripts/trace_dkv.py (build with -DDML_TRACE)
    const bool tr0 = p.trace && blockIdx.x == 0 && lane == 0 && warp == 0;
    const bool tr3 = p.trace && blockIdx.x == 0 && lane == 0 && warp == 3;
#else
    constexpr bool tr0 = false, tr3 = false;
#endif
    const int nit = ntiles - t_begin;
    for (int t = 0; t < nit; ++t) {                      // t counts this CTA's tiles
      const int st = t % kStages, buf = t & 1;
      mbar_wait_relaxed(bar(kBarInFull + st), (t / kStages) & 1);
      mbar_wait_relaxed(bar(kBarSFull + buf), (t >> 1) & 1);
      tc_fence_after();
      if (tr0) p.trace[t * 8 + 1] = clock64();
      if (tr3) p.trace[t * 8 + 3] = clock64();
      const uint32_t rowa = sbase + kOffRow + st * kRowStride + half * 64;
      const uint32_t tS = tmem + lane_off + buf * 128 + head * 32 + half * 16;
      // segment of this thread's first position and the boundaries after it -> which variant the whole warp takes
      int seg_first, seg_last, mode;
      float3 bnd = make_float3(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000));
      float2 clo = make_float2(0.f, 0.f), chi = clo;
      const float x_first = cpb_x(lds_f32(rowa) - g_j), x_last = cpb_x(lds_f32(rowa + 15 * 4) - g_j);
      if (SL.staged) {
        // x only grows from tile to tile: step on from the segment the last tile ended in (run.seg)
        int sf = run.seg;
        if (sf < 0) { int c; lookup_seg(L, SL, x_first, c, sf); }
        else { while (x_first >= lds_f32(SL.bp + (uint32_t)sf * 4u)) ++sf; }
        seg_first = sf;
        bnd.x = lds_f32(SL.bp + (uint32_t)sf * 4u);
        bnd.y = lds_f32(SL.bp + (uint32_t)sf * 4u + 4u);
        if (!__any_sync(0xffffffffu, x_last >= bnd.y)) {
          seg_last = sf + (x_last >= bnd.x ? 1 : 0);
          clo = lds_f32x2(SL.coef + (uint32_t)sf * 16u + (head ? 8u : 0u));
          chi = lds_f32x2(SL.coef + (uint32_t)sf * 16u + 16u + (head ? 8u : 0u));
          mode = __any_sync(0xffffffffu, x_last >= bnd.x) ? 0 : 3;
        } else {
          bnd.z = lds_f32(SL.bp + (uint32_t)sf * 4u + 8u);
          int sl = sf;
          while (x_last >= lds_f32(SL.bp + (uint32_t)sl * 4u)) ++sl;
          seg_last = sl;
          mode = __any_sync(0xffffffffu, sl - sf > 3) ? 2 : 1;
        }
      } else {
        int c0, c1;
        lookup2<true, true>(L, x_first, c0, seg_first);
        lookup2<true, true>(L, x_last, c1, seg_last);
        mode = 2;
      }
      h16* const dsp = ds_row ? ds_row + t * kBI : nullptr;      // ds_row already points at this CTA's first tile
#define DML_DKV_SWEEP(M, E) dkv_sweep<M, E>(L, SL, tS, rowa, head, g_j, kvld, sc2, seg_first, seg_last, bnd, clo, chi, dgacc, run, ssum, dsp, tr0 ? p.trace + t * 8 + 0 : nullptr)
      if (!key_masked) {
        if (mode == 3) DML_DKV_SWEEP(false, 3);
        else if (mode == 0) DML_DKV_SWEEP(false, 0);
        else if (mode == 1) DML_DKV_SWEEP(false, 1);
        else DML_DKV_SWEEP(false, 2);
      } else {
        if (mode == 3) DML_DKV_SWEEP(true, 3);
        else if (mode == 0) DML_DKV_SWEEP(true, 0);
        else if (mode == 1) DML_DKV_SWEEP(true, 1);
        else DML_DKV_SWEEP(true, 2);
      }
#undef DML_DKV_SWEEP
      if (tr0) p.trace[t * 8 + 2] = clock64();
      if (tr3) p.trace[t * 8 + 4] = clock64();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(kBarPFull + buf));
    }
    seg_flush_warp(ssum, true, run.seg, run.a, run.b, head, lane);
    if (kvld) atomicAdd(p.dg + (size_t)(b * G + grp) * p.n_kv + gj, dgacc * inv_s);

    // ---- drain dV / dK ----
    mbar_wait(bar(kBarAcc), 0);
    tc_fence_after();
    const float ksc = p.scale * inv_s;
    const int ldg = p.H * kD;
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2) {        // dV h0, dV h1, dK h0, dK h1: 64 columns each; this warp: its head's two, 32 columns of each
      const int acc = a2 * 2 + head;
      const int h = head;
      float* dst = (acc < 2 ? p.dv : p.dk) + ((size_t)b * p.n_kv + gj) * ldg + (h0 + h) * kD + half * 32;
      const float f = acc < 2 ? inv_s : ksc;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t a[16];
        tmem_ld16(tmem + lane_off + 256 + acc * 64 + half * 32 + c * 16, a);
        tmem_ld_wait(a);
        if (kvld) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 v = make_float4(__uint_as_float(a[e]) * f, __uint_as_float(a[e + 1]) * f, __uint_as_float(a[e + 2]) * f,
                                         __uint_as_float(a[e + 3]) * f);
            float4* o = reinterpret_cast<float4*>(dst + c * 16 + e);
            if (whole) *o = v;
            else atomicAdd(o, v);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEwWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// =================================================================================================================
// dQ = dS K as a streaming GEMM over the dS^T workspace
// =================================================================================================================
namespace dqg {
constexpr int kBM = 128, kBKey = 64, kStages = 4, kThreads = 192;   // 4 drain warps, TMA producer, MMA issuer
constexpr uint32_t kTileA = kBKey * 64 * 2;       // 8 KB: 64 keys x 64 queries of dS^T (one 128-byte swizzle span wide)
constexpr uint32_t kTileK = kBKey * kD * 2;       // 8 KB
constexpr uint32_t kStageBytes = 4 * kTileA + 2 * kTileK;   // per head: two query halves of dS^T; K0 K1
constexpr uint32_t kOffBar = kStages * kStageBytes;
constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAcc = 2 * kStages, kNumBars = kBarAcc + 1;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemPtr + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr uint32_t kIdesc = idesc_f16(128, 64, true, true);    // A = dS^T tile ([key][query]: MN-major), B = K tile ([key][d]: MN-major)
// MN-major A of 128 queries = two 64-query spans kTileA bytes apart (leading-dimension byte offset), 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t desc_a(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(kTileA >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
}  // namespace dqg

__global__ void __launch_bounds__(dqg::kThreads, 1)
deform_attn_dq_gemm_kernel(const __grid_constant__ CUtensorMap mds, const __grid_constant__ CUtensorMap mk, const BwdParams p) {
  using namespace dqg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  const int i0 = blockIdx.x * kBM, grp = blockIdx.y, b = blockIdx.z;
  const int h0 = grp * 2;
  const int nsteps = p.n_kv_pad / kBKey;
  auto bar = [&](int i) { return sbase + kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kOffTmemPtr);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarFull + s), 1); mbar_init(bar(kBarEmpty + s), 1); }
    mbar_init(bar(kBarAcc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmemPtr), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---- TMA producer ----
    if (lane == 0) {
      for (int j = 0; j < nsteps; ++j) {
        const int st = j % kStages;
        mbar_wait(bar(kBarEmpty + st), ((j / kStages) & 1) ^ 1);
        const uint32_t dst = sbase + st * kStageBytes;
        mbar_expect_tx(bar(kBarFull + st), kStageBytes);
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(dst + (2 * h) * kTileA, &mds, bar(kBarFull + st), i0, j * kBKey, b * p.H + h0 + h);
          tma_load_3d(dst + (2 * h + 1) * kTileA, &mds, bar(kBarFull + st), i0 + 64, j * kBKey, b * p.H + h0 + h);
          tma_load_3d(dst + 4 * kTileA + h * kTileK, &mk, bar(kBarFull + st), (h0 + h) * kD, j * kBKey, b);
        }
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer ----
    const bool leader = elect_one();
    for (int j = 0; j < nsteps; ++j) {
      const int st = j % kStages;
      mbar_wait(bar(kBarFull + st), (j / kStages) & 1);
      tc_fence_after();
      const uint32_t src = sbase + st * kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t da = desc_a(src + (2 * h) * kTileA), db = smem_desc(src + 4 * kTileA + h * kTileK);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tmem + h * 64, da + 128 * k, db + 128 * k, kIdesc, (j > 0) || (k > 0), leader);
      }
      tc_commit(bar(kBarEmpty + st), leader);
    }
    tc_commit(bar(kBarAcc), leader);
  } else {
    // ---- drain: dQ / s -> global ----
    const int gi = i0 + warp * 32 + lane;
    const float inv_s = __ldg(p.dscale + 1);
    mbar_wait(bar(kBarAcc), 0);
    tc_fence_after();
    float* ob = p.dq + ((size_t)b * p.n + gi) * (p.H * kD) + h0 * kD;
    const uint32_t tb = tmem + (((uint32_t)warp * 32u) << 16);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t a[16];
      tmem_ld16(tb + c * 16, a);
      tmem_ld_wait(a);
      if (gi < p.n) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(ob + c * 16 + e) =
              make_float4(__uint_as_float(a[e]) * inv_s, __uint_as_float(a[e + 1]) * inv_s,
                          __uint_as_float(a[e + 2]) * inv_s, __uint_as_float(a[e + 3]) * inv_s);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
  }
}

}  // namespace tc
}  // namespace dml

// ---- host: work list of the dK/dV kernel (see the launch in dml_deform_attn_bwd_tc) ----------------------------------
static bool dkv_worklist_applies(int items, int ntiles, int nsm) {
  return items <= nsm && 2 * nsm <= dml::tc::kDkvWorkMax && ntiles >= 96 && items < 65536 && ntiles < 65536;
}
// Cuts the items' query-tile ranges into pieces of about equal estimated cost (one cut per SM, never inside the first or
// last 24 tiles of what it would split) and lists them longest first; returns the number of pieces.
static int plan_dkv_worklist(int B, int G, int n, int n_kv, int nsm, dml::tc::DkvWorkList& wl) {
  using namespace dml;
  using namespace dml::tc;
  const int nkb = cdiv(n_kv, dkvk::kBK), items = nkb * G * B, ntiles = cdiv(n, dkvk::kBI);
  const double kDense = 1.45, half = 0.16 * n / dkvk::kBI;
  auto cost = [&](int item, int t) {
    const double tdiag = ((item % nkb) * dkvk::kBK + 0.5 * dkvk::kBK) * ((double)n / n_kv) / dkvk::kBI;
    return fabs(t - tdiag) < half ? kDense : 1.0;
  };
  double total = 0.0;
  for (int i = 0; i < items; ++i)
    for (int t = 0; t < ntiles; ++t) total += cost(i, t);
  const double target = total / nsm;
  double acc = 0.0, pc[kDkvWorkMax];
  int k = 1, np = 0;
  for (int i = 0; i < items && np < kDkvWorkMax; ++i) {
    int t0 = 0;
    double c = 0.0;
    for (int t = 0; t < ntiles; ++t) {
      const double ct = cost(i, t);
      acc += ct; c += ct;
      if (acc >= k * target - 1e-9 && k < nsm) {
        ++k;
        // cut here unless it would leave a sliver (each piece pays a ~10 us prologue): then the item boundary or the
        // previous cut stands in for it
        if (t + 1 - t0 >= 24 && ntiles - (t + 1) >= 24 && np < kDkvWorkMax - 1) {
          wl.e[np] = DkvWork{(uint16_t)i, (uint16_t)t0, (uint16_t)(t + 1), 0}; pc[np++] = c;
          t0 = t + 1; c = 0.0;
        }
      }
    }
    wl.e[np] = DkvWork{(uint16_t)i, (uint16_t)t0, (uint16_t)ntiles, 0}; pc[np++] = c;
  }
  for (int a = 1; a < np; ++a) {      // insertion sort, longest first (np <= 320)
    const DkvWork w = wl.e[a];
    const double c = pc[a];
    int q = a - 1;
    for (; q >= 0 && pc[q] < c; --q) { wl.e[q + 1] = wl.e[q]; pc[q + 1] = pc[q]; }
    wl.e[q + 1] = w; pc[q + 1] = c;
  }
  return np;
}

extern "C" {

#ifdef DML_TEST_KNOBS
/* debug: device buffer of long long[4 * 2 * ntiles] that the next dQ launches fill with clock64() stamps (NULL = off) */
int dml_debug_set_trace(void* buf) {
  dml::tc::g_trace = (long long*)buf;
  return 0;
}
#endif

/* bytes of the optional dS^T workspace of dml_deform_attn_bwd_tc: fp16 [(B H), ceil128(n_kv), ceil32(n)] */
size_t dml_deform_attn_bwd_ws_bytes(int B, int H, int n, int n_kv) {
  if (B <= 0 || H <= 0 || n <= 0 || n_kv <= 0) return 0;
  return (size_t)B * H * (size_t)(dml::cdiv(n_kv, 128) * 128) * (size_t)(dml::cdiv(n, 32) * 32) * 2;
}

/* test aid (host only, no device needed): the dK/dV work list for a problem shape on a device with `nsm` SMs as
 * (item, t0, t1) triples in launch order; returns the number of pieces (0: one CTA per item), at most cap are written */
int dml_debug_dkv_worklist(int B, int H, int n, int n_kv, int nsm, int* out, int cap) {
  using namespace dml;
  using namespace dml::tc;
  if (B <= 0 || H < 2 || n <= 0 || n_kv <= 0 || nsm <= 0 || !out) return DML_EINVAL;
  const int G = H / 2, items = cdiv(n_kv, dkvk::kBK) * G * B, ntiles = cdiv(n, dkvk::kBI);
  if (!dkv_worklist_applies(items, ntiles, nsm)) return 0;
  DkvWorkList wl;
  const int np = plan_dkv_worklist(B, G, n, n_kv, nsm, wl);
  for (int i = 0; i < np && i < cap; ++i) {
    out[3 * i] = wl.e[i].item; out[3 * i + 1] = wl.e[i].t0; out[3 * i + 2] = wl.e[i].t1;
  }
  return np;
}

#ifdef DML_TEST_KNOBS
/* test knob: tables with at least limit - 2 segments are treated as too large for the shared-memory segment arrays of the
 * dK/dV kernel (its general per-position path); limit <= 0 restores the default */
int dml_debug_set_seg_limit(int limit) {
  dml::tc::g_seg_limit = limit > 0 ? limit : dml::tc::kSegSmem;
  return 0;
}
#endif

int dml_deform_attn_bwd_tc(const void* q, const void* k, const void* v, const float* g, const void* table,
                           const void* out, const void* d_out, const float* lse, int B, int H, int dim_head, int n,
                           int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo, int heads_per_group, float scale,
                           const float* dscale, float* dsum_ws, float* dq, float* dk, float* dv, float* dg,
                           float* segsum, void* ds_ws, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  DML_CHECK_ARG(q && k && v && g && table && out && d_out && lse && dscale && dsum_ws && dk && dv && dg && segsum);
  // dq == NULL with a workspace: stop after the dK / dV / dS^T stage; the caller finishes with dml_deform_attn_dq_from_ds (on
  // another stream if it likes: everything that only needs dK / dV / dg can then run next to the HBM-bound dQ GEMM)
  DML_CHECK_ARG(dq || ds_ws);
  if (((uintptr_t)ds_ws) & 15) return DML_EINVAL;
  DML_CHECK_ARG(B > 0 && H > 0 && n > 0 && n_kv > 0 && n_seq >= n);
  if (dim_head != kD || heads_per_group != 2 || (H & 1)) return DML_EUNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return DML_EINVAL;
  if (ldq < H * kD || ldk < H * kD || ldv < H * kD || ldo != H * kD) return DML_EINVAL;
  if ((((uintptr_t)q) | ((uintptr_t)k) | ((uintptr_t)v) | ((uintptr_t)d_out) | ((uintptr_t)dq) | ((uintptr_t)dk) |
       ((uintptr_t)dv)) & 15)
    return DML_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mq128, mdo128, mk32, mv32, mq32, mdo32, mk128, mv128;
  int rc;
  if ((rc = make_map(&mq128, q, B, n, ldq, 128)) || (rc = make_map(&mdo128, d_out, B, n, ldo, 128)) ||
      (rc = make_map(&mk32, k, B, n_kv, ldk, 32)) || (rc = make_map(&mv32, v, B, n_kv, ldv, 32)) ||
      (rc = make_map(&mq32, q, B, n, ldq, 32)) || (rc = make_map(&mdo32, d_out, B, n, ldo, 32)) ||
      (rc = make_map(&mk128, k, B, n_kv, ldk, 128)) || (rc = make_map(&mv128, v, B, n_kv, ldv, 128)))
    return rc;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(deform_attn_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dqk::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(deform_attn_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkvk::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(deform_attn_dq_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dqg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  const int G = H / 2;
  cudaError_t e = cudaMemsetAsync(dg, 0, sizeof(float) * (size_t)B * G * n_kv, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(segsum, 0, sizeof(float) * 4 * kCpbSegMax, st);
  if (e != cudaSuccess) return (int)e;
  BwdParams p{};
  p.g = g; p.table = (const uint32_t*)table; p.lse = lse; p.dsum = dsum_ws; p.dscale = dscale;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dg = dg; p.segsum = segsum;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.n_seq = n_seq; p.scale = scale;
#ifdef DML_TEST_KNOBS
  p.trace = dml::tc::g_trace;
  p.seg_limit = dml::tc::g_seg_limit;
#else
  p.trace = nullptr;
  p.seg_limit = dml::tc::kSegSmem;
#endif
  p.ds_ws = (h16*)ds_ws; p.n_pad = cdiv(n, 32) * 32; p.n_kv_pad = cdiv(n_kv, 128) * 128;
  CUtensorMap mds, mk64;
  if (ds_ws) {
    if ((rc = make_map(&mds, ds_ws, B * H, p.n_kv_pad, p.n_pad, 64)) || (rc = make_map(&mk64, k, B, n_kv, ldk, 64))) return rc;
  }
  const int rows = B * n;
  bwd_prep_kernel<<<min(cdiv(rows, 8), 148 * 8), 256, 0, st>>>((const float*)out, (const h16*)d_out, B, n, H, ldo, dsum_ws);
  {
    // Decomposition of the dK/dV kernel.  One CTA per item (batch, head pair, 128-key block) is right when the items fill
    // the SMs several times over.  With few items (128 on 148 SMs at the north-star size) the items' query-tile ranges
    // are cut into pieces of about equal estimated cost - one cut per SM, so a piece never spans two items - and listed
    // longest first; the hardware hands CTAs to SMs in index order, which makes this the longest-processing-time rule
    // (a simulation with the cost model below: makespan 527 vs 599 tile units unsplit, 502 ideal).  Cost model: tiles
    // within 0.16 n positions of the item's diagonal - where the table segments are dense - count 1.45, each piece pays
    // a prologue of 12 tiles.  DML_B200_DKV_QSPLIT=<q> forces q equal parts per item instead (tuning aid; 1 = unsplit).
    int nsm = 0;       // per device (the current one), not cached process-wide; it is part of the work-list cache key below
    {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    }
    static int forced = -1;
    if (forced < 0) {
      const char* ev = getenv("DML_B200_DKV_QSPLIT");
      forced = ev ? atoi(ev) : 0;
    }
    const int nkb = cdiv(n_kv, dkvk::kBK), items = nkb * G * B, ntiles = cdiv(n, dkvk::kBI);
    // the list only depends on the problem shape: each host thread keeps the last one it built (65 k cost evaluations
    // would otherwise sit between the prep kernel and this launch on every eager call)
    thread_local DkvWorkList wl;
    thread_local long long wl_key[5] = {-1, -1, -1, -1, -1};
    const long long key[5] = {B, G, n, n_kv, nsm};
    const bool cached = forced <= 0 && key[0] == wl_key[0] && key[1] == wl_key[1] && key[2] == wl_key[2] && key[3] == wl_key[3] && key[4] == wl_key[4];
    int ncta = items;
    p.qsplit = 1;
    if (cached) {
      if (wl.n > 0) ncta = wl.n;
    } else if (forced > 0) {
      wl.n = 0;
      wl_key[0] = -1;
      p.qsplit = ntiles / forced >= 1 ? forced : 1;
      ncta = items * p.qsplit;
    } else if (!dkv_worklist_applies(items, ntiles, nsm)) {
      wl.n = 0;
      for (int q = 0; q < 5; ++q) wl_key[q] = key[q];
    } else {
      const int np = plan_dkv_worklist(B, G, n, n_kv, nsm, wl);
      wl.n = np;
      ncta = np;
      for (int q = 0; q < 5; ++q) wl_key[q] = key[q];
    }
    if (ncta != items) {
      if ((e = cudaMemsetAsync(dk, 0, sizeof(float) * (size_t)B * n_kv * H * kD, st)) != cudaSuccess) return (int)e;
      if ((e = cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)B * n_kv * H * kD, st)) != cudaSuccess) return (int)e;
    }
    deform_attn_dkv_tc_kernel<<<ncta, dkvk::kThreads, dkvk::kSmemBytes, st>>>(mq32, mdo32, mk128, mv128, p, wl);
  }
  if (ds_ws) {
    if (dq) deform_attn_dq_gemm_kernel<<<dim3(cdiv(n, dqg::kBM), G, B), dqg::kThreads, dqg::kSmemBytes, st>>>(mds, mk64, p);
  } else
    deform_attn_dq_tc_kernel<<<dim3(cdiv(n, dqk::kGroups * dqk::kBM), G, B), dqk::kThreads, dqk::kSmemBytes, st>>>(mq128, mdo128, mk32, mv32, p);
  DML_RETURN_LAUNCH();
}

/* The last stage of the workspace backward on its own: dq [B, n, H*64] (fp32) = (1/s) * dS . K from a dS^T workspace laid
 * out as dml_deform_attn_bwd_tc writes it (fp16 [(B H), ceil128(n_kv), ceil32(n)], times the loss scale s = dscale[0]).
 * A streaming GEMM: reads the workspace once - the HBM-bound kernel of the path (bench.py times it against the copy
 * bandwidth).                                                                                                          */
int dml_deform_attn_dq_from_ds(const void* ds_ws, const void* k, const float* dscale, int B, int H, int dim_head, int n,
                               int n_kv, int ldk, float* dq, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  DML_CHECK_ARG(ds_ws && k && dscale && dq && B > 0 && H > 0 && n > 0 && n_kv > 0);
  if (dim_head != kD || (H & 1)) return DML_EUNSUPPORTED;
  if ((ldk % 8) || ldk < H * kD) return DML_EINVAL;
  if ((((uintptr_t)ds_ws) | ((uintptr_t)k) | ((uintptr_t)dq)) & 15) return DML_EINVAL;
  BwdParams p{};
  p.dscale = dscale; p.dq = dq; p.B = B; p.H = H; p.n = n; p.n_kv = n_kv;
  p.ds_ws = (h16*)ds_ws; p.n_pad = cdiv(n, 32) * 32; p.n_kv_pad = cdiv(n_kv, 128) * 128;
  CUtensorMap mds, mk64;
  int rc;
  if ((rc = make_map(&mds, ds_ws, B * H, p.n_kv_pad, p.n_pad, 64)) || (rc = make_map(&mk64, k, B, n_kv, ldk, 64))) return rc;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(deform_attn_dq_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dqg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  deform_attn_dq_gemm_kernel<<<dim3(cdiv(n, dqg::kBM), H / 2, B), dqg::kThreads, dqg::kSmemBytes, (cudaStream_t)stream>>>(mds, mk64, p);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
