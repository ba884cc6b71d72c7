// Fused deformable cross-attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Reference maths (DeformableAttention1D.py:203-231 + CPB :84-102), per head h of offset group grp = h / 2:
//     S[i,j] = scale * q_h[i,:].k_h[j,:] + bias_{h % 2}(x_ij),   x_ij = sign(p) log2(|p| + 1),  p = seq_i - g[grp, j]
//     P = softmax_j(S);  out_h[i,:] = sum_j P[i,j] v_h[j,:]
//
// One CTA = one (batch, offset group, 256 consecutive queries): the TWO heads of the group are processed together,
// so the position x_ij, its log and its table cell are evaluated once per (i, j) and shared by both heads.
//
//   warp 16   TMA producer: Q tiles once, then a 4-stage ring of {K_h0, K_h1, V_h0, V_h1} 32-key tiles
//             (cp.async.bulk.tensor, 128-byte swizzle) + the 32 sampling positions g of the tile
//   warps 17,18 MMA issuers, one per query group (uniform datapath, one elected lane): S = Q K^T (SS form, operands in
//             shared memory) and O += P V (TS form: P is read from TMEM, V from shared memory as an MN-major operand);
//             warp 17 owns the TMEM allocation
//   warps 0-7 softmax group 0 = queries [i0, i0+128);  warps 8-15 softmax group 1 = queries [i0+128, i0+256).
//             Inside a group warp w takes TMEM lanes 32 (w & 3).. (query rows) and the key half (w >> 2) & 1 of every
//             32-key tile: TMEM lane t holds query row t's S row and O row, and a row is shared by exactly two threads
//             (one per key half, in two warps).  Each keeps a partial row sum; the row maximum of the raw S tile is
//             exchanged through shared memory behind a 64-thread named barrier, after which both threads derive the
//             same softmax reference.  Four softmax warps per scheduler instead of two hide the table-lookup / MUFU /
//             TMEM latencies of each other.  The 32 lanes of a warp look up nearly the same table cell (consecutive
//             queries) -> broadcast shared-memory reads.
//
// TMEM (512 columns x 128 lanes, fp32): group g at column 256 g:  two S buffers of 64 columns (S_h0 32 | S_h1 32) at 0 and
// 64, O_h0 [128,192), O_h1 [192,256).  S of key tile j+1 is computed into the other buffer while the softmax warps work
// on tile j, so the MMA hand-off latency is hidden.  P (fp16) overwrites S in place: the 16 fp32 columns of key chunk c
// become 8 columns of P_hi and 8 columns of P_lo (P = P_hi + P_lo, 22 significant bits), consumed by two K=16 MMAs.
//
// Online softmax without a correction pass in the common case: before the main sweep a cheap sweep over the raw S
// tile gives an upper bound of the row maximum (max S + an upper bound of the piecewise-linear bias over the tile's
// position window); the running reference m is only raised when that bound exceeds it by more than 2^8, and only then
// are the O rows rescaled in TMEM (after waiting for the group's previous O += P V to complete).
#include <math.h>

#include "../../include/dml_b200.h"
#include "tc_common.cuh"

namespace dml {
namespace tc {

constexpr int kD = 64;            // head dim
constexpr int kBM = 128;          // query rows per softmax group (= TMEM lanes)
constexpr int kGroups = 2;        // softmax groups per CTA
constexpr int kBN = 32;           // keys per tile
constexpr int kStages = 4;
constexpr int kSoftWarps = 16;      // 2 groups x 4 lane quarters x 2 key halves
constexpr int kThreads = 32 * (kSoftWarps + 3);   // + TMA producer + one MMA-issuing warp per group
constexpr uint32_t kTileQ = kBM * kD * 2;   // 16384 B
constexpr uint32_t kTileKV = kBN * kD * 2;  // 4096 B
constexpr uint32_t kStageBytes = 4 * kTileKV;
constexpr float kRaise = 8.0f;    // raise the softmax reference only when the bound exceeds it by 2^8

// shared-memory map (dynamic, base aligned to 1024 B)
constexpr uint32_t kOffQ = 0;                                        // [group][head] 128x64 fp16
constexpr uint32_t kOffKV = kOffQ + kGroups * 2 * kTileQ;            // [stage]{K0,K1,V0,V1} 64x64 fp16
constexpr uint32_t kOffG = kOffKV + kStages * kStageBytes;           // [stage][32 g + gmin + gmax + pad] floats
constexpr uint32_t kGStride = 40 * 4;
constexpr uint32_t kOffRec = kOffG + kStages * kGStride;             // bias table image (tc_common.cuh), 16-B aligned
constexpr uint32_t kOffPair = kOffRec + kTabSmemBytes;               // [slot 2][group 2][row 128][key half 2] float2: row-max exchange
constexpr uint32_t kPairBytes = 2 * kGroups * kBM * 2 * 8;
constexpr uint32_t kOffPairBh = kOffPair + kPairBytes;               // [slot 2][group 2][row 128] float4: bias bound + flagged-cell count
constexpr uint32_t kPairBhBytes = 2 * kGroups * kBM * 16;
constexpr uint32_t kOffBar = kOffPairBh + kPairBhBytes;              // mbarriers (8 B each)
constexpr int kBarQ = 0, kBarKvFull = 1, kBarKvEmpty = kBarKvFull + kStages, kBarSFull = kBarKvEmpty + kStages,   // [group][buffer]
              kBarPFull = kBarSFull + 2 * kGroups, kBarPvDone = kBarPFull + 2 * kGroups, kBarOFinal = kBarPvDone + kGroups,
              kNumBars = kBarOFinal + kGroups;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemPtr + 16 + 1024;             // + slack for the 1024-B alignment
static_assert(kOffRec % 16 == 0 && kOffPair % 8 == 0 && kOffPairBh % 16 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct Params {
  const float* g;        // [(B G), n_kv] normalised sampling positions
  const uint32_t* table; // cpb table (device)
  float* o;              // fp32 [B, n, ldo]
  float* lse;            // [B, H, n] (log2 domain)
  int B, H, n, n_kv, ldo, n_seq;
  float scale;
  int n_full, n_half;    // per (bag, offset group): 256-query CTAs, then 128-query CTAs (launched after every 256-query one)
};

constexpr uint32_t kIdescS = idesc_f16(128, kBN, false, false);   // S = Q K^T: A, B K-major
constexpr uint32_t kIdescPV = idesc_f16(128, 64, false, true);    // O += P V: A in TMEM, B (= V) MN-major

// One key half (16 keys) of a 32-key tile of one query row, both heads, from registers: sa / sb = the row's 16 S values
// of head 0 / 1 -> P = exp2(S sc2 + bias - m) split into fp16 hi/lo pairs written over the same S columns (S_h0 at tS,
// S_h1 at tS + 32; tS already points at the half's first column); l0/l1 accumulate the partial row sums.  kMasked: keys
// >= jrem (counted inside the half) are padding; kDirty: the row's position window touches a table cell holding >= 2
// breakpoints.  gsa = shared address of the half's 16 sampling positions.
template <bool kMasked, bool kDirty>
__device__ __forceinline__ void sweep2(const Lookup& L, uint32_t tS, uint32_t gsa, const uint32_t (&sa)[16],
                                       const uint32_t (&sb)[16], float s_i, float sc2, float m0, float m1, int jrem,
                                       float& l0, float& l1) {
  float gq[16];
#pragma unroll
  for (int e = 0; e < 16; e += 4) {
    const float4 t = lds_f32x4(gsa + (uint32_t)e * 4);
    gq[e] = t.x; gq[e + 1] = t.y; gq[e + 2] = t.z; gq[e + 3] = t.w;
  }
  uint32_t w0[16], w1[16];   // [0,8) = P_hi pairs, [8,16) = P_lo pairs
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    float v0[2], v1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float x = cpb_x(s_i - gq[e + u]);
      int cdummy, sdummy;
      const float4 t = lookup2<kDirty, false>(L, x, cdummy, sdummy);
      v0[u] = ex2(fmaf(__uint_as_float(sa[e + u]), sc2, fmaf(t.x, x, t.y)) - m0);
      v1[u] = ex2(fmaf(__uint_as_float(sb[e + u]), sc2, fmaf(t.z, x, t.w)) - m1);
      if (kMasked && e + u >= jrem) { v0[u] = 0.f; v1[u] = 0.f; }
      l0 += v0[u];
      l1 += v1[u];
    }
    split_f16(v0[0], v0[1], w0[e >> 1], w0[8 + (e >> 1)]);
    split_f16(v1[0], v1[1], w1[e >> 1], w1[8 + (e >> 1)]);
  }
  tmem_st16(tS, w0);
  tmem_st16(tS + 32, w1);
}

// 64-thread named barrier of the two warps that share a lane quarter of a group (ids 1..8; id 0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
deform_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                          const __grid_constant__ CUtensorMap mv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
  // blockIdx.x enumerates the 256-query CTAs of every (bag, group) first, then the 128-query ones (one softmax group active):
  // the hardware hands CTAs out in index order, so the short ones fill the last, partial wave
  const int G = p.H / 2;
  int i0, grp, b, ngroups;
  {
    const int nf = p.n_full * G * p.B;
    int x = blockIdx.x, per = p.n_full;
    ngroups = kGroups;
    if (x >= nf) { x -= nf; per = p.n_half; ngroups = 1; }
    const int blk = x % per, gb = x / per;
    i0 = ngroups == kGroups ? blk * (kGroups * kBM) : p.n_full * (kGroups * kBM) + blk * kBM;
    grp = gb % G;
    b = gb / G;
  }
  const int ntiles = cdiv(p.n_kv, kBN);
  auto bar = [&](int i) { return sbase + kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + kOffTmemPtr);

  // ---- one-time setup ----
  if (tid == 0) {
    mbar_init(bar(kBarQ), 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(kBarKvFull + s), 32); mbar_init(bar(kBarKvEmpty + s), ngroups); }
    for (int g = 0; g < 2 * kGroups; ++g) { mbar_init(bar(kBarSFull + g), 1); mbar_init(bar(kBarPFull + g), 2 * kBM); }
    for (int g = 0; g < kGroups; ++g) { mbar_init(bar(kBarPvDone + g), 1); mbar_init(bar(kBarOFinal + g), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kSoftWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const Lookup L = tab_stage(sgen + kOffRec, sbase + kOffRec, p.table, tid, kThreads);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tab_finish(sgen + kOffRec, lane);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kSoftWarps) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_expect_tx(bar(kBarQ), ngroups * 2 * kTileQ);
      for (int g = 0; g < ngroups; ++g)
        for (int h = 0; h < 2; ++h)
          tma_load_3d(sbase + kOffQ + (g * 2 + h) * kTileQ, &mq, bar(kBarQ), (grp * 2 + h) * kD, i0 + g * kBM, b);
    }
    const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      mbar_wait_relaxed(bar(kBarKvEmpty + st), ((j / kStages) & 1) ^ 1);
      const uint32_t dst = sbase + kOffKV + st * kStageBytes;
      if (lane == 0) {
        mbar_expect_tx_only(bar(kBarKvFull + st), kStageBytes);
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(dst + h * kTileKV, &mk, bar(kBarKvFull + st), (grp * 2 + h) * kD, j * kBN, b);
          tma_load_3d(dst + (2 + h) * kTileKV, &mv, bar(kBarKvFull + st), (grp * 2 + h) * kD, j * kBN, b);
        }
      }
      float* gs = reinterpret_cast<float*>(sgen + kOffG + st * kGStride);
      // |p| <= 1 + |g| must stay inside the table domain (log2(|p| + 1) < X): clamp, so that no cell index can leave the table
      const float gb_max = tab_gmax(p.table);
      const float g0 = fminf(fmaxf(__ldg(gb + min(j * kBN + lane, p.n_kv - 1)), -gb_max), gb_max);
      gs[lane] = g0;
      const float gmn = -warp_max(-g0), gmx = warp_max(g0);
      if (lane == 0) { gs[32] = gmn; gs[33] = gmx; }
      mbar_arrive(bar(kBarKvFull + st));                     // 32 arrivals (release) + the TMA bytes complete the phase
    }
  } else if (warp > kSoftWarps + ngroups || (warp < kSoftWarps && (warp >> 3) >= ngroups)) {
    // the second query group of a 128-query CTA: its MMA warp and its eight softmax warps have nothing to do
  } else if (warp > kSoftWarps) {
    // =========================== MMA issuers: one warp per query group ===========================
    // All 32 lanes run the loop (uniform operands); one elected lane executes each tcgen05 instruction.
    const int g = warp - (kSoftWarps + 1);
    const bool leader = elect_one();
    mbar_wait(bar(kBarQ), 0);
    auto issue_s = [&](int j) {
      const int st = j % kStages, buf = j & 1;
      const uint32_t kv = sbase + kOffKV + st * kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t da = smem_desc(sbase + kOffQ + (g * 2 + h) * kTileQ), db = smem_desc(kv + h * kTileKV);
        const uint32_t d = tmem + g * 256 + buf * 64 + h * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(d, da + 2 * k, db + 2 * k, kIdescS, k > 0, leader);
      }
      tc_commit(bar(kBarSFull + g * 2 + buf), leader);
    };
    mbar_wait(bar(kBarKvFull + 0), 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages, buf = j & 1;
      if (j + 1 < ntiles) {            // S of the next tile -> the other buffer (its P(j-1) was consumed by the PV MMAs of j-1)
        mbar_wait(bar(kBarKvFull + (j + 1) % kStages), ((j + 1) / kStages) & 1);
        tc_fence_after();
        issue_s(j + 1);
      }
      const uint32_t kv = sbase + kOffKV + st * kStageBytes;
      mbar_wait_relaxed(bar(kBarPFull + g * 2 + buf), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t db = smem_desc(kv + (2 + h) * kTileKV);
        const uint32_t d = tmem + g * 256 + 128 + h * 64, a = tmem + g * 256 + buf * 64 + h * 32;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          mma_ts(d, a + 16 * k, db + 128 * k, kIdescPV, (j > 0) || (k > 0), leader);       // P_hi chunk k
          mma_ts(d, a + 16 * k + 8, db + 128 * k, kIdescPV, 1, leader);                    // P_lo chunk k
        }
      }
      tc_commit(bar(kBarKvEmpty + st), leader);        // this group's share: the stage is free once both groups' MMAs are done
      tc_commit(bar(kBarPvDone + g), leader);          // O holds the contributions of tiles 0..j
    }
    tc_commit(bar(kBarOFinal + g), leader);            // one-shot: every MMA of the group has completed
  } else {
    // =========================== softmax groups ===========================
    const int g = warp >> 3;                        // group
    const int kh = (warp >> 2) & 1;                 // key half of every tile
    const int row = (warp & 3) * 32 + lane;         // TMEM lane = query row inside the group's tile
    const int gi = i0 + g * kBM + row;
    const int pair_id = 1 + g * 4 + (warp & 3);     // named barrier shared with the warp of the other key half
    const uint32_t tbase = tmem + g * 256 + (((uint32_t)(warp & 3) * 32u) << 16);
    const float amax0 = __uint_as_float(__ldg(p.table + 6)), amax1 = __uint_as_float(__ldg(p.table + 7));
    const float s_i = seq_pos(min(gi, p.n_seq - 1), p.n_seq);     // rows past the end: any in-domain position
    const float sc2 = p.scale * kLog2e;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // m: common to both halves; l: this half's partial sum
    // exchange slots: [slot][group][row][half] float2
    auto pair_slot = [&](int slot, int half) { return sbase + kOffPair + (uint32_t)(((slot * kGroups + g) * kBM + row) * 2 + half) * 8u; };

    for (int j = 0; j < ntiles; ++j) {
      const int st = j % kStages;
      const uint32_t gsa = sbase + kOffG + st * kGStride;
      const int buf = j & 1;
      const uint32_t tS = tbase + buf * 64 + kh * 16;
      mbar_wait(bar(kBarKvFull + st), (j / kStages) & 1);       // g tile visible to this thread
      mbar_wait(bar(kBarSFull + g * 2 + buf), (j >> 1) & 1);    // S(j) landed
      tc_fence_after();
      const int jrem = p.n_kv - j * kBN - kh * 16;              // valid keys in this half of the tile (may be <= 0)

      // ---- this half of the row's S tile (16 keys x 2 heads) into registers, once; its maximum goes to the partner ----
      uint32_t sa[16], sb[16];
      tmem_ld16(tS, sa);
      tmem_ld16(tS + 32, sb);
      tmem_ld_fence();
      reg_fence(sa); reg_fence(sb);
      float r0 = -INFINITY, r1 = -INFINITY;
      if (jrem >= 16) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          r0 = fmaxf(r0, __uint_as_float(sa[e]));
          r1 = fmaxf(r1, __uint_as_float(sb[e]));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (e < jrem) {
            r0 = fmaxf(r0, __uint_as_float(sa[e]));
            r1 = fmaxf(r1, __uint_as_float(sb[e]));
          }
      }
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pair_slot(buf, kh)), "f"(r0), "f"(r1) : "memory");
      // the bias bound over the tile's position window is the same for both halves of the row: the kh = 0 thread works
      // it out and hands it over with the maximum (the two warps share a scheduler: the issue slots go to the partner)
      const uint32_t bh_slot = sbase + kOffPairBh + (uint32_t)((buf * kGroups + g) * kBM + row) * 16u;
      float bh0 = 0.f, bh1 = 0.f;
      int ndirty = 0;
      if (kh == 0) {
        const float xlo = cpb_x(s_i - lds_f32(gsa + 33 * 4)), xhi = cpb_x(s_i - lds_f32(gsa + 32 * 4));
        int clo, chi, sdummy;
        const float4 e = lookup2<true, false>(L, xlo, clo, sdummy), f = lookup2<true, false>(L, xhi, chi, sdummy);
        const float half = 0.5f * (xhi - xlo) + 1e-6f;
        bh0 = fmaxf(fmaf(e.x, xlo, e.y), fmaf(f.x, xhi, f.y)) + amax0 * half;
        bh1 = fmaxf(fmaf(e.z, xlo, e.w), fmaf(f.z, xhi, f.w)) + amax1 * half;
        ndirty = tab_dirty_between(L, clo, chi);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bh_slot), "f"(bh0), "f"(bh1), "f"(__int_as_float(ndirty)), "f"(0.f) : "memory");
      }
      pair_sync(pair_id);
      {
        const float2 o = lds_f32x2(pair_slot(buf, kh ^ 1));
        r0 = fmaxf(r0, o.x);
        r1 = fmaxf(r1, o.y);
        if (kh == 1) {
          const float4 v = lds_f32x4(bh_slot);
          bh0 = v.x; bh1 = v.y; ndirty = __float_as_int(v.z);
        }
      }
      // from here on both threads of the row hold identical r, bh, m: they take the same decisions
      const float ub0 = fmaf(r0, sc2, bh0), ub1 = fmaf(r1, sc2, bh1);
      const bool raise0 = ub0 > m0 + kRaise, raise1 = ub1 > m1 + kRaise;
      if (__any_sync(0xffffffffu, raise0 || raise1)) {
        const float f0 = raise0 ? ex2(m0 - ub0) : 1.0f, f1 = raise1 ? ex2(m1 - ub1) : 1.0f;   // ex2(-inf) = 0 on the first tile
        if (raise0) { m0 = ub0; l0 *= f0; }
        if (raise1) { m1 = ub1; l1 *= f1; }
        if (j > 0) {                                           // rescale this warp's share of the O rows in TMEM
          mbar_wait(bar(kBarPvDone + g), (j - 1) & 1);          // O += P V of tile j-1 has completed
          tc_fence_after();
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) {
            const int c = c2 * 2 + kh;                         // 16-column chunks of O_h0 (0..3) and O_h1 (4..7), alternating
            uint32_t a[16];
            tmem_ld16(tbase + 128 + c * 16, a);
            tmem_ld_wait(a);
            const float f = c < 4 ? f0 : f1;
#pragma unroll
            for (int e = 0; e < 16; ++e) a[e] = __float_as_uint(__uint_as_float(a[e]) * f);
            tmem_st16(tbase + 128 + c * 16, a);
          }
        }
      }

      // ---- sweep 2: P = exp2(S sc2 + bias - m), in place ----
      const bool dirty = __any_sync(0xffffffffu, ndirty != 0);
      const uint32_t gh = gsa + kh * 64;
      if (jrem >= 16) {
        if (!dirty) sweep2<false, false>(L, tS, gh, sa, sb, s_i, sc2, m0, m1, jrem, l0, l1);
        else sweep2<false, true>(L, tS, gh, sa, sb, s_i, sc2, m0, m1, jrem, l0, l1);
      } else {
        sweep2<true, true>(L, tS, gh, sa, sb, s_i, sc2, m0, m1, jrem, l0, l1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(kBarPFull + g * 2 + buf));
    }

    // ---- epilogue: the two partial row sums -> l; O / l -> global, log-sum-exp ----
    // (a one-shot barrier: the per-tile kBarPvDone may still be several phases behind here, and a parity wait is only
    // meaningful when the waiter is at most one phase ahead)
    {
      const int slot = ntiles & 1;                             // the slot the last tile did not use
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(pair_slot(slot, kh)), "f"(l0), "f"(l1) : "memory");
      pair_sync(pair_id);
      const float2 o = lds_f32x2(pair_slot(slot, kh ^ 1));
      l0 += o.x;
      l1 += o.y;
    }
    mbar_wait(bar(kBarOFinal + g), 0);
    tc_fence_after();
    const int h0 = grp * 2;
    if (gi < p.n && kh == 0) {
      float* lb = p.lse + ((size_t)b * p.H + h0) * p.n + gi;
      lb[0] = m0 + __log2f(l0);
      lb[p.n] = m1 + __log2f(l1);
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    float* ob = p.o + ((size_t)b * p.n + gi) * p.ldo + h0 * kD;
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      const int c = c2 * 2 + kh;
      uint32_t a[16];
      tmem_ld16(tbase + 128 + c * 16, a);      // warp-collective: every lane takes part, stores are predicated
      tmem_ld_wait(a);
      const float f = c < 4 ? inv0 : inv1;
      if (gi < p.n) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(ob + c * 16 + e) =
              make_float4(__uint_as_float(a[e]) * f, __uint_as_float(a[e + 1]) * f, __uint_as_float(a[e + 2]) * f,
                          __uint_as_float(a[e + 3]) * f);
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace tc
}  // namespace dml

extern "C" {

int dml_deform_attn_fwd_tc(const void* q, const void* k, const void* v, const float* g, const void* table, int B,
                           int H, int dim_head, int n, int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo,
                           int heads_per_group, float scale, void* out, float* lse, void* stream) {
  return dml_deform_attn_fwd_tc_split(q, k, v, g, table, B, H, dim_head, n, n_kv, n_seq, ldq, ldk, ldv, ldo, heads_per_group, scale,
                                      out, lse, 0, stream);
}

int dml_deform_attn_fwd_tc_split(const void* q, const void* k, const void* v, const float* g, const void* table, int B,
                                 int H, int dim_head, int n, int n_kv, int n_seq, int ldq, int ldk, int ldv, int ldo,
                                 int heads_per_group, float scale, void* out, float* lse, int half_blocks, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  DML_CHECK_ARG(q && k && v && g && table && out && lse && B > 0 && H > 0 && n > 0 && n_kv > 0 && n_seq >= n && half_blocks >= 0);
  if (dim_head != kD || heads_per_group != 2 || (H & 1)) return DML_EUNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 4)) return DML_EINVAL;
  if (ldq < H * kD || ldk < H * kD || ldv < H * kD || ldo < H * kD) return DML_EINVAL;
  if ((((uintptr_t)q) | ((uintptr_t)k) | ((uintptr_t)v) | ((uintptr_t)out)) & 15) return DML_EINVAL;
  CUtensorMap mq, mk, mv;
  int rc = make_map(&mq, q, B, n, ldq, kBM);
  if (rc) return rc;
  rc = make_map(&mk, k, B, n_kv, ldk, kBN);
  if (rc) return rc;
  rc = make_map(&mv, v, B, n_kv, ldv, kBN);
  if (rc) return rc;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(deform_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  Params p{};
  p.g = g; p.table = (const uint32_t*)table; p.o = (float*)out; p.lse = lse;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.ldo = ldo; p.n_seq = n_seq; p.scale = scale;
  // 129 row groups of 128 queries at n = 16 385: 64 CTAs of two groups + one of one group per (bag, offset group); the last
  // half_blocks two-group blocks of every (bag, group) run as twice as many one-group CTAs, launched after all two-group ones
  const int rg = cdiv(n, kBM);
  const int tail = half_blocks;
  p.n_full = max(rg / kGroups - tail, 0);
  p.n_half = rg - kGroups * p.n_full;
  const long long ncta = (long long)(p.n_full + p.n_half) * (H / 2) * B;
  if (ncta > 0x7fffffffLL) return DML_EUNSUPPORTED;
  deform_attn_fwd_tc_kernel<<<(unsigned)ncta, kThreads, kSmemBytes, (cudaStream_t)stream>>>(mq, mk, mv, p);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
