// dml_pgemm: batched GEMM on tcgen05 / TMEM / TMA with fp32-class accuracy from bf16 operand PAIRS (sm_100a).
//
// Every fp32 tensor x that enters a contraction of the path is held as two bf16 planes, x ~= hi + lo (hi = bf16(x),
// lo = bf16(x - hi): 16 significant bits, the fp32 exponent range - no scale factors, no absmax pass), and a product is
// accumulated in fp32 in TMEM as hi.hi + hi.lo + lo.hi (three tcgen05.mma per k-step; an operand that is exact in bf16,
// such as the bf16 bag, has one plane and costs two).  Against the reference (fp32 torch) this gives 6-9e-6 relative error
// through the whole NystromAttention forward and backward, pseudo-inverse recurrence included (TF32: 2e-3, outside the
// 1e-3 bar) - measured on the reference's own goldens, DESIGN.md section 3.
//
// One kernel serves every contraction of NystromAttention (models/NystromAttention.py:89,122-125,138-140,150), TransMIL's
// fc1 (models/mil.py:229) and the projections of DeformCrossTransMIL / DeformCrossAttention1D
// (models/DeformCrossTransMIL.py:100,111; models/DeformableAttention1D.py:175,199,233) and all of their gradients:
//   * either operand may be K-major (memory [rows][K]) or MN-major (memory [K][rows]) - the instruction descriptor's
//     major bits - so C = A B^T, A B, A^T B all read the row-major tensors as they are (no transposes, no copies);
//   * row / k offsets with TMA zero fill give the FRONT padding of NystromAttention (:79-85) without a padded copy;
//   * two batch dimensions with arbitrary element strides address head slices inside a fused [n, 3 H d] qkv buffer;
//   * the epilogue (TMEM lane = output row, one thread per row) fuses: alpha (with a second factor for a leading column
//     range: q * scale), per-column bias, residual add, accumulate, ReLU, "diag I - C" (the pinv polynomial terms), the row
//     softmax of the landmark similarity products and its backward, and writes fp32 and / or the bf16 pair of the result
//     and / or fp16 (the operand type of the fused attention kernels), plus an optional |C| maximum;
//   * split-K (fp32 reductions into a zeroed C) for the token-reduction products (weight gradients, attn3 @ v).
//
// Persistent CTAs (one per SM) walk the 128 x BN output tiles (BN = 64 / 128 / 256) with two TMEM accumulator buffers, so the
// epilogue of a tile overlaps the loads and MMAs of the next; warp 8 = TMA producer (ring of {A planes, B planes} 64-wide k-blocks,
// 128-byte swizzle), warp 9 = MMA issuer + TMEM owner, warps 0-7 = epilogue (a row's columns are split between two threads,
// 32-column groups, vector loads / stores; one CTA per SM, so the epilogue's own parallelism is what hides its latencies).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "../../include/dml_b200.h"
#include "tc_common.cuh"

namespace dml {
namespace tc {
namespace pg {

constexpr int kBM = 128, kBK = 64, kThreads = 32 * 10;   // 8 epilogue warps, TMA producer, MMA issuer
constexpr uint32_t kTileA = kBM * kBK * 2;   // 16 KB

template <int BN>
struct Cfg {
  static constexpr int kStages = BN == 64 ? 4 : (BN == 128 ? 3 : 2);
  static constexpr uint32_t kTileB = BN * kBK * 2;
  static constexpr uint32_t kStageBytes = 2 * kTileA + 2 * kTileB;
  static constexpr uint32_t kOffBar = kStages * kStageBytes;
  static constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAccFull = 2 * kStages, kBarAccEmpty = 2 * kStages + 2,
                       kNumBars = 2 * kStages + 4;
  static constexpr uint32_t kTmemCols = 2 * BN;                  // two accumulator buffers
  static constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr uint32_t kOffXch = kOffTmemPtr + 16;          // float[2][128]
  static constexpr uint32_t kSmemBytes = kOffXch + 1024 + 1024;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

struct Params {
  int M, N, K;
  int nb_inner, splits;
  int a_layout, b_layout, a_planes, b_planes;
  int a_row_off, a_k_off, b_row_off, b_k_off;
  int a_bi, a_bo, b_bi, b_bo;                 // 1: the operand is indexed by that batch dimension, 0: shared
  float alpha, alpha2;
  int ncol_split;
  const float* alpha_dev;
  const float* bias; long long bias_bi, bias_bo;
  int relu, use_diag;
  float diag;
  const float* resid; int ldr; long long r_bi, r_bo; float resid_scale;
  int accumulate;
  float* c; int ldc; long long c_bi, c_bo;
  bf16* pair; int ldp; long long p_bi, p_bo, p_plane;
  h16* half_out; int ldh; long long h_bi, h_bo;
  const float* half_scale_dev;
  uint32_t* absmax;
  int softmax;
  const bf16* aux; int ldx; long long x_bi, x_bo, x_plane;
  int tiles_m, tiles_n, total_tiles;
  int vec_ok;      // every output / side tensor allows 16-byte accesses at 4- (fp32) / 8- (16-bit) element column granularity
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// MN-major operand: 64-element spans of the M / N index 8 KB apart (leading-dimension byte offset), 8-row k groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// One GEMM problem on this CTA: every role walks the problem's tiles (blockIdx.x, + gridDim.x, ...).  `it` (k-blocks through
// the shared-memory ring) and `ti` (tiles through the two TMEM accumulator buffers) are the calling thread's running
// counters: they continue across problems when a kernel chains several of them.
template <int BN>
__device__ __forceinline__ void run_problem(const CUtensorMap* ma, const CUtensorMap* mb, const Params& p, uint32_t sbase,
                                            uint8_t* sgen, uint32_t tmem, int warp, int lane, int& it, int& ti) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  const int nk_all = cdiv(p.K, kBK);
  const int kb_per = cdiv(nk_all, p.splits);
  // persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; tile index = (z * tiles_n + n_tile) * tiles_m + m_tile, so
  // the CTAs running at the same time share the B tile.  Two TMEM accumulator buffers: the MMAs of tile i + 1 run while the
  // epilogue warps drain tile i.
  struct Tile { int m0, n0, bi, bo, kb0, nk; };
  auto decode = [&](int t) {
    Tile T;
    const int mt = t % p.tiles_m, r = t / p.tiles_m;
    const int nt = r % p.tiles_n, z = r / p.tiles_n;
    const int split = z % p.splits, batch = z / p.splits;
    T.m0 = mt * kBM; T.n0 = nt * BN; T.bi = batch % p.nb_inner; T.bo = batch / p.nb_inner;
    T.kb0 = split * kb_per;
    T.nk = max(min(nk_all, T.kb0 + kb_per) - T.kb0, 0);
    return T;
  };
  auto bar = [&](int i) { return sbase + C::kOffBar + 8u * i; };
  if (warp == 8) {
    // ---- TMA producer ----
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)p.a_planes * kTileA + (uint32_t)p.b_planes * C::kTileB;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const Tile T = decode(t);
        const int m0 = T.m0, n0 = T.n0;
        const int abi = p.a_bi ? T.bi : 0, abo = p.a_bo ? T.bo : 0, bbi = p.b_bi ? T.bi : 0, bbo = p.b_bo ? T.bo : 0;
        for (int kb = 0; kb < T.nk; ++kb, ++it) {
          const int st = it % kStages;
          mbar_wait(bar(C::kBarEmpty + st), ((it / kStages) & 1) ^ 1);
          const uint32_t dst = sbase + st * C::kStageBytes;
          const uint32_t fb = bar(C::kBarFull + st);
          mbar_expect_tx(fb, bytes);
          const int k0 = (T.kb0 + kb) * kBK;
          for (int pl = 0; pl < p.a_planes; ++pl) {
            const uint32_t d = dst + pl * kTileA;
            if (p.a_layout == 0) {
              tma_load_5d(d, ma, fb, k0 + p.a_k_off, m0 + p.a_row_off, abi, abo, pl);
            } else {
              tma_load_5d(d, ma, fb, m0 + p.a_row_off, k0 + p.a_k_off, abi, abo, pl);
              tma_load_5d(d + 8192, ma, fb, m0 + 64 + p.a_row_off, k0 + p.a_k_off, abi, abo, pl);
            }
          }
          for (int pl = 0; pl < p.b_planes; ++pl) {
            const uint32_t d = dst + 2 * kTileA + pl * C::kTileB;
            if (p.b_layout == 0) {
              tma_load_5d(d, mb, fb, k0 + p.b_k_off, n0 + p.b_row_off, bbi, bbo, pl);
            } else {
#pragma unroll
              for (int s = 0; s < BN / 64; ++s)
                tma_load_5d(d + s * 8192, mb, fb, n0 + 64 * s + p.b_row_off, k0 + p.b_k_off, bbi, bbo, pl);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---- MMA issuer (uniform datapath, one elected lane) ----
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16(128, BN, p.a_layout != 0, p.b_layout != 0);
    const uint32_t ka = p.a_layout ? 128u : 2u, kbs = p.b_layout ? 128u : 2u;   // descriptor advance per 16 k (16-byte units)
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
      const Tile T = decode(t);
      const int buf = ti & 1;
      mbar_wait(bar(C::kBarAccEmpty + buf), ((ti >> 1) & 1) ^ 1);        // the epilogue has drained this buffer
      tc_fence_after();
      const uint32_t acc = tmem + buf * BN;
      for (int kb = 0; kb < T.nk; ++kb, ++it) {
        const int st = it % kStages;
        mbar_wait(bar(C::kBarFull + st), (it / kStages) & 1);
        tc_fence_after();
        const uint32_t base = sbase + st * C::kStageBytes;
        const uint64_t a0 = p.a_layout ? desc_mn(base) : smem_desc(base);
        const uint64_t a1 = p.a_layout ? desc_mn(base + kTileA) : smem_desc(base + kTileA);
        const uint64_t b0 = p.b_layout ? desc_mn(base + 2 * kTileA) : smem_desc(base + 2 * kTileA);
        const uint64_t b1 = p.b_layout ? desc_mn(base + 2 * kTileA + C::kTileB) : smem_desc(base + 2 * kTileA + C::kTileB);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          mma_ss(acc, a0 + ka * k, b0 + kbs * k, idesc, (kb > 0) || (k > 0), leader);
          if (p.b_planes > 1) mma_ss(acc, a0 + ka * k, b1 + kbs * k, idesc, 1, leader);
          if (p.a_planes > 1) mma_ss(acc, a1 + ka * k, b0 + kbs * k, idesc, 1, leader);
        }
        tc_commit(bar(C::kBarEmpty + st), leader);
      }
      tc_commit(bar(C::kBarAccFull + buf), leader);
    }
  } else {
    // ---- epilogue: 8 warps; TMEM lane = output row, the row's columns split between two threads (warps w and w + 4) ----
    const int quarter = warp & 3, half = warp >> 2;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
    const Tile T = decode(t);
    const int m0 = T.m0, n0 = T.n0, bi = T.bi, bo = T.bo, nk = T.nk;
    const int buf = ti & 1;
    const int row = m0 + quarter * 32 + lane;
    const bool rv = row < p.M;
    float alpha = p.alpha;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    const float alpha_q = alpha * p.alpha2;
    const size_t rowl = (size_t)(rv ? row : 0);
    const float* biasp = p.bias ? p.bias + (size_t)bo * p.bias_bo + (size_t)bi * p.bias_bi : nullptr;
    const float* resp = p.resid ? p.resid + (size_t)bo * p.r_bo + (size_t)bi * p.r_bi + rowl * p.ldr : nullptr;
    float* cp = p.c ? p.c + (size_t)bo * p.c_bo + (size_t)bi * p.c_bi + rowl * p.ldc : nullptr;
    bf16* pp = p.pair ? p.pair + (size_t)bo * p.p_bo + (size_t)bi * p.p_bi + rowl * p.ldp : nullptr;
    h16* hp = p.half_out ? p.half_out + (size_t)bo * p.h_bo + (size_t)bi * p.h_bi + rowl * p.ldh : nullptr;
    const bf16* xp = p.aux ? p.aux + (size_t)bo * p.x_bo + (size_t)bi * p.x_bi + rowl * p.ldx : nullptr;
    const float hscale = (p.half_out && p.half_scale_dev) ? __ldg(p.half_scale_dev) : 1.0f;
    const uint32_t tb = tmem + buf * BN + (((uint32_t)quarter * 32u) << 16);
    float* xch = reinterpret_cast<float*>(sgen + C::kOffXch);          // [2 halves][128 rows] exchange of row statistics
    mbar_wait(bar(C::kBarAccFull + buf), (ti >> 1) & 1);
    tc_fence_after();
    // this thread's column groups (32 wide): [g0, g1) of the tile
    constexpr int kGroups = BN / 32;
    const int g0 = half * (kGroups / 2), g1 = g0 + kGroups / 2;
    // accumulator group -> alpha * acc (no other stage)
    auto load_acc = [&](int g, float (&v)[32], float a_lo, float a_hi) {
      uint32_t r0[16], r1[16];
      if (nk > 0) {
        tmem_ld16(tb + g * 32, r0);
        tmem_ld16(tb + g * 32 + 16, r1);
        tmem_ld_wait2(r0, r1);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r0[e] = r1[e] = 0u;
      }
      const int col0 = n0 + g * 32;
      if (col0 + 32 <= p.ncol_split) {
#pragma unroll
        for (int e = 0; e < 16; ++e) { v[e] = __uint_as_float(r0[e]) * a_lo; v[16 + e] = __uint_as_float(r1[e]) * a_lo; }
      } else if (col0 >= p.ncol_split) {
#pragma unroll
        for (int e = 0; e < 16; ++e) { v[e] = __uint_as_float(r0[e]) * a_hi; v[16 + e] = __uint_as_float(r1[e]) * a_hi; }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          v[e] = __uint_as_float(r0[e]) * (col0 + e < p.ncol_split ? a_lo : a_hi);
          v[16 + e] = __uint_as_float(r1[e]) * (col0 + 16 + e < p.ncol_split ? a_lo : a_hi);
        }
      }
    };
    // the additive / pointwise stages on a group whose 32 columns all exist (vector accesses; host checked the alignment)
    auto stages_full = [&](int col0, float (&v)[32]) {
      if (biasp) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(biasp + col0 + e));
          v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
        }
      }
      if (resp && rv) {
        const float rs = p.resid_scale;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(resp + col0 + e));
          v[e] = fmaf(rs, t.x, v[e]); v[e + 1] = fmaf(rs, t.y, v[e + 1]); v[e + 2] = fmaf(rs, t.z, v[e + 2]); v[e + 3] = fmaf(rs, t.w, v[e + 3]);
        }
      }
      if (p.accumulate && rv) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(cp + col0 + e);
          v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
        }
      }
      if (p.relu) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      if (p.use_diag) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = (row == col0 + e ? p.diag : 0.f) - v[e];
      }
    };
    auto stages_edge = [&](int col0, float (&v)[32]) {       // per-element version for a ragged last group / unaligned tensors
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = col0 + e;
        float x = v[e];
        if (rv && col < p.N) {
          if (biasp) x += __ldg(biasp + col);
          if (resp) x = fmaf(p.resid_scale, __ldg(resp + col), x);
          if (p.accumulate) x += cp[col];
        }
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.use_diag) x = (row == col ? p.diag : 0.f) - x;
        v[e] = x;
      }
    };
    auto store_full = [&](int col0, const float (&v)[32]) {
      if (cp) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(cp + col0 + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
      if (pp) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        bf16* ph = pp + col0;
        bf16* pl = ph + p.p_plane;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          *reinterpret_cast<uint4*>(ph + 2 * e) = make_uint4(hi[e], hi[e + 1], hi[e + 2], hi[e + 3]);
          *reinterpret_cast<uint4*>(pl + 2 * e) = make_uint4(lo[e], lo[e + 1], lo[e + 2], lo[e + 3]);
        }
      }
      if (hp) {
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = pack_f16(v[2 * e] * hscale, v[2 * e + 1] * hscale);
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<uint4*>(hp + col0 + 2 * e) = make_uint4(w[e], w[e + 1], w[e + 2], w[e + 3]);
      }
    };
    auto store_edge = [&](int col0, const float (&v)[32]) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = col0 + e;
        if (col < p.N) {
          if (cp) cp[col] = v[e];
          if (pp) {
            const bf16 h = __float2bfloat16_rn(v[e]);
            pp[col] = h;
            pp[p.p_plane + col] = __float2bfloat16_rn(v[e] - __bfloat162float(h));
          }
          if (hp) hp[col] = __float2half_rn(v[e] * hscale);
        }
      }
    };
    auto load_aux = [&](int col0, bool full, float (&a)[32]) {
      if (full && rv) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const uint4 h = *reinterpret_cast<const uint4*>(xp + col0 + e), l = *reinterpret_cast<const uint4*>(xp + p.x_plane + col0 + e);
          a[e] = bf16_lo_f(h.x) + bf16_lo_f(l.x); a[e + 1] = bf16_hi_f(h.x) + bf16_hi_f(l.x);
          a[e + 2] = bf16_lo_f(h.y) + bf16_lo_f(l.y); a[e + 3] = bf16_hi_f(h.y) + bf16_hi_f(l.y);
          a[e + 4] = bf16_lo_f(h.z) + bf16_lo_f(l.z); a[e + 5] = bf16_hi_f(h.z) + bf16_hi_f(l.z);
          a[e + 6] = bf16_lo_f(h.w) + bf16_lo_f(l.w); a[e + 7] = bf16_hi_f(h.w) + bf16_hi_f(l.w);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          a[e] = (rv && col0 + e < p.N) ? __bfloat162float(xp[col0 + e]) + __bfloat162float(xp[p.x_plane + col0 + e]) : 0.f;
      }
    };
    // combine a per-half row statistic across the two threads of a row (all 8 epilogue warps take part)
    auto exchange = [&](float mine, bool is_max) -> float {
      const int rl = quarter * 32 + lane;
      xch[half * 128 + rl] = mine;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float other = xch[(half ^ 1) * 128 + rl];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      return is_max ? fmaxf(mine, other) : mine + other;
    };
    const bool vec = p.vec_ok != 0;
    float amax = 0.f;
    if (p.splits > 1) {
      // split-K: reduce alpha * partial into the zeroed C
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        const int col0 = n0 + g * 32;
        if (col0 >= p.N) break;
        float v[32];
        load_acc(g, v, alpha_q, alpha);
        if (rv && nk > 0) {
          if (vec && col0 + 32 <= p.N) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) atomicAdd(reinterpret_cast<float4*>(cp + col0 + e), make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col0 + e < p.N) atomicAdd(cp + col0 + e, v[e]);
          }
        }
      }
    } else if (p.softmax == 0) {
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        const int col0 = n0 + g * 32;
        if (col0 >= p.N) break;
        const bool full = vec && col0 + 32 <= p.N;
        float v[32];
        load_acc(g, v, alpha_q, alpha);
        if (full) stages_full(col0, v); else stages_edge(col0, v);
        if (p.absmax) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (rv && col0 + e < p.N) amax = fmaxf(amax, fabsf(v[e]));
        }
        if (rv) {
          if (full) store_full(col0, v); else store_edge(col0, v);
        }
      }
    } else if (p.softmax == 1) {
      // row softmax over the N columns of this tile (one column tile; value = alpha * acc): exponent in the log2 domain
      const float al2 = alpha * kLog2e;
      float mx = -INFINITY;
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        if (g * 32 >= p.N) break;
        float v[32];
        load_acc(g, v, al2, al2);
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (g * 32 + e < p.N) mx = fmaxf(mx, v[e]);
      }
      mx = exchange(mx, true);
      float sum = 0.f;
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        if (g * 32 >= p.N) break;
        float v[32];
        load_acc(g, v, al2, al2);
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (g * 32 + e < p.N) sum += ex2(v[e] - mx);
      }
      sum = exchange(sum, false);
      const float inv = 1.0f / sum;
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        const int col0 = g * 32;
        if (col0 >= p.N) break;
        float v[32];
        load_acc(g, v, al2, al2);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = ex2(v[e] - mx) * inv;
        if (rv) {
          if (vec && col0 + 32 <= p.N) store_full(col0, v); else store_edge(col0, v);
        }
      }
    } else {
      // softmax backward: value = alpha * acc = dA, aux = A (pair): dS = A * (dA - sum_j dA_j A_j)
      float dot = 0.f;
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        const int col0 = g * 32;
        if (col0 >= p.N) break;
        float v[32], a[32];
        load_acc(g, v, alpha, alpha);
        load_aux(col0, vec && col0 + 32 <= p.N, a);
#pragma unroll
        for (int e = 0; e < 32; ++e) dot = fmaf(v[e], a[e], dot);      // aux is zero beyond N
      }
      dot = exchange(dot, false);
#pragma unroll 1
      for (int g = g0; g < g1; ++g) {
        const int col0 = g * 32;
        if (col0 >= p.N) break;
        const bool full = vec && col0 + 32 <= p.N;
        float v[32], a[32];
        load_acc(g, v, alpha, alpha);
        load_aux(col0, full, a);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = a[e] * (v[e] - dot);
        if (rv) {
          if (full) store_full(col0, v); else store_edge(col0, v);
        }
      }
    }
    if (p.absmax) {
      amax = warp_max(amax);
      if (lane == 0 && amax > 0.f) atomicMax(p.absmax, __float_as_uint(amax));
    }
    // this warp's TMEM reads of the tile are complete (every tcgen05.ld above was waited for): hand the buffer back
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(C::kBarAccEmpty + buf));
    }   // tiles
  }

}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
pgemm_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, const Params p) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + C::kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + C::kOffTmemPtr);
  // Programmatic dependent launch: the next kernel of the stream may start its own prologue (barrier init, TMEM allocation,
  // descriptor fetch) on idle SMs while this grid is still running - chains of small dependent products (the pseudo-inverse
  // recurrence: 32 CTAs each) are bound by exactly that per-launch latency.  Nothing below touches global memory before
  // griddepcontrol.wait, which returns once the preceding grid has completed and its writes are visible.
  if (threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(C::kBarFull + s), 1); mbar_init(bar(C::kBarEmpty + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar(C::kBarAccFull + b), 1); mbar_init(bar(C::kBarAccEmpty + b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::kOffTmemPtr), "r"(C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  int it = 0, ti = 0;
  run_problem<BN>(&ma, &mb, p, sbase, sgen, tmem, warp, lane, it, ti);

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::kTmemCols));
  }
}

// Several dependent GEMM problems in ONE cooperative launch (the products of the pseudo-inverse recurrence,
// models/NystromAttention.py:31-33, and of its adjoint): problem i + 1 reads what problem i wrote, so the CTAs meet at a grid
// barrier between problems.  Global writes of the epilogue (generic proxy) are made visible to the TMA loads (async proxy) of
// the other CTAs by a proxy fence on both sides of the barrier.  Every problem has at most gridDim.x tiles.
constexpr int kMaxChain = 24;
struct ChainParams {
  int count;
  CUtensorMap maps[2 * kMaxChain];
  Params p[kMaxChain];
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
pgemm_chain_kernel(const __grid_constant__ ChainParams cp) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + C::kOffBar + 8u * i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + C::kOffTmemPtr);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(C::kBarFull + s), 1); mbar_init(bar(C::kBarEmpty + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar(C::kBarAccFull + b), 1); mbar_init(bar(C::kBarAccEmpty + b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::kOffTmemPtr), "r"(C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  int it = 0, ti = 0;
  for (int i = 0; i < cp.count; ++i) {
    run_problem<BN>(&cp.maps[2 * i], &cp.maps[2 * i + 1], cp.p[i], sbase, sgen, tmem, warp, lane, it, ti);
    if (i + 1 < cp.count) {
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      grid.sync();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::kTmemCols));
  }
}

// fp32 [rows, cols] (row stride ld) -> bf16 pair planes [rows, ldp]; optional per-call factor
__global__ void __launch_bounds__(256)
pair_from_f32_kernel(const float* __restrict__ x, long long rows, int cols, int ld, float mult, bf16* __restrict__ out, int ldp,
                     long long plane) {
  const int cols4 = (cols + 3) >> 2;
  const long long total = rows * cols4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    float v[4];
    const float* src = x + r * ld + c;
    if (c + 4 <= cols && ((ld & 3) == 0) && ((((uintptr_t)x) & 15) == 0)) {
      const float4 f = *reinterpret_cast<const float4*>(src);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = c + e < cols ? src[e] : 0.f;
    }
    uint32_t hi[2], lo[2];
    split_bf16x2(v[0] * mult, v[1] * mult, hi[0], lo[0]);
    split_bf16x2(v[2] * mult, v[3] * mult, hi[1], lo[1]);
    bf16* oh = out + r * ldp + c;
    bf16* ol = oh + plane;
    if (c + 4 <= ldp) {
      *reinterpret_cast<uint2*>(oh) = make_uint2(hi[0], hi[1]);
      *reinterpret_cast<uint2*>(ol) = make_uint2(lo[0], lo[1]);
    } else {
      for (int e = 0; e < 4 && c + e < ldp; ++e) {
        reinterpret_cast<uint16_t*>(oh)[e] = (uint16_t)((e & 1) ? (hi[e >> 1] >> 16) : (hi[e >> 1] & 0xffffu));
        reinterpret_cast<uint16_t*>(ol)[e] = (uint16_t)((e & 1) ? (lo[e >> 1] >> 16) : (lo[e >> 1] & 0xffffu));
      }
    }
  }
}

// ReLU backward fused with the pair conversion: gm[r, c] = act[r, c] > 0 ? g[r, c] : 0, as fp32 (gm may alias g) and as a bf16 pair
__global__ void __launch_bounds__(256)
relu_mask_pair_kernel(const float* __restrict__ g, float* __restrict__ gm, const float* __restrict__ act, long long rows, int cols,
                      int ldg, int lda, bf16* __restrict__ out, int ldp, long long plane) {
  const int cols4 = cols >> 2;
  const long long total = rows * cols4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    float4 v = *reinterpret_cast<const float4*>(g + r * ldg + c);
    const float4 a = *reinterpret_cast<const float4*>(act + r * lda + c);
    v.x = a.x > 0.f ? v.x : 0.f; v.y = a.y > 0.f ? v.y : 0.f; v.z = a.z > 0.f ? v.z : 0.f; v.w = a.w > 0.f ? v.w : 0.f;
    *reinterpret_cast<float4*>(gm + r * ldg + c) = v;
    bf16* oh = out + r * ldp + c;
    store_pair4(oh, oh + plane, v.x, v.y, v.z, v.w);
  }
}

// (s, 1 / s) with s the power of two that puts max|t| into (4, 8] (s = 1 for an all-zero tensor, |log2 s| <= 60), from the bit
// pattern of max|t| that the GEMM epilogue left (`absmax`): one thread, exponent arithmetic only
__global__ void loss_scale_kernel(const unsigned* __restrict__ amax_bits, float* __restrict__ out) {
  const unsigned bits = *amax_bits & 0x7fffffffu;
  const float a = __uint_as_float(bits);
  float s = 1.0f;
  if (a > 0.f) {
    // a = m 2^e, 1 <= m < 2 (denormals: frexp-style through the float multiply below):  floor(log2(8 / a)) = 3 - e - (m > 1)
    int e, ex;
    const float m = frexpf(fmaxf(a, 1e-30f), &e);      // a = m 2^e with 0.5 <= m < 1
    ex = 3 - (e - 1) - (m > 0.5f ? 1 : 0);
    ex = max(-60, min(60, ex));
    s = ldexpf(1.0f, ex);
  }
  out[0] = s;
  out[1] = 1.0f / s;
}

// x * (*scale) -> fp16, 8 elements per thread (the loss-scaled dO operand of the attention backward)
__global__ void __launch_bounds__(256)
scale_to_half_kernel(const float* __restrict__ x, const float* __restrict__ scale, long long n8, h16* __restrict__ out) {
  const float s = __ldg(scale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pack_f16(a.x * s, a.y * s), pack_f16(a.z * s, a.w * s), pack_f16(b.x * s, b.y * s),
                                                        pack_f16(b.z * s, b.w * s));
  }
}

// column sums of a [rows, cols] fp32 matrix (bias gradients): 32 x 8 threads per CTA over row chunks, atomics into zeroed out
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int cols, int ld, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long rpb = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + ty; r < r1; r += 8) s += x[r * ld + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tx];
    if (c < cols) atomicAdd(out + c, s);
  }
}

static int make_map5(CUtensorMap* m, const dml_pg_operand& o, int K, int nb_inner, int nb_outer, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DML_EUNSUPPORTED;
  const int planes = o.plane_stride ? 2 : 1;
  const cuuint64_t nbi = o.bs_inner ? (cuuint64_t)nb_inner : 1, nbo = o.bs_outer ? (cuuint64_t)nb_outer : 1;
  const cuuint64_t k_mem = (cuuint64_t)(o.k_mem > 0 ? o.k_mem : K);
  cuuint64_t dims[5];
  cuuint32_t box[5] = {64, 64, 1, 1, 1};
  if (o.layout == 0) { dims[0] = k_mem; dims[1] = (cuuint64_t)o.rows; box[1] = (cuuint32_t)box_rows; }
  else { dims[0] = (cuuint64_t)o.rows; dims[1] = k_mem; }
  dims[2] = nbi; dims[3] = nbo; dims[4] = (cuuint64_t)planes;
  // strides of dimensions that are never stepped (extent 1) only have to be legal
  const cuuint64_t dflt = (cuuint64_t)o.ld * 2 * dims[1];
  cuuint64_t strides[4] = {(cuuint64_t)o.ld * 2, o.bs_inner ? (cuuint64_t)o.bs_inner * 2 : dflt,
                           o.bs_outer ? (cuuint64_t)o.bs_outer * 2 : dflt, o.plane_stride ? (cuuint64_t)o.plane_stride * 2 : dflt};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(o.base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DML_OK : DML_EINVAL;
}

static bool operand_ok(const dml_pg_operand& o) {
  if (!o.base || (((uintptr_t)o.base) & 15)) return false;
  if (o.ld <= 0 || (o.ld % 8) || o.rows <= 0) return false;
  if ((o.bs_inner % 8) || (o.bs_outer % 8) || (o.plane_stride % 8)) return false;
  if (o.bs_inner < 0 || o.bs_outer < 0 || o.plane_stride < 0) return false;
  return o.layout == 0 || o.layout == 1;
}

}  // namespace pg
}  // namespace tc
}  // namespace dml

extern "C" {

// validate one problem, encode its tensor maps and fill the kernel parameters; *bn_out = the tile width the problem needs
static int sm_count_cached() {      // per call, from the current device (no process-global state)
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)
    nsm = 148;
  return nsm;
}

static int prepare_problem(const dml_pgemm_args* a, dml::tc::pg::Params& p, CUtensorMap* ma, CUtensorMap* mb, int* bn_out,
                           bool narrow_small) {
  using namespace dml;
  using namespace dml::tc;
  using namespace dml::tc::pg;
  DML_CHECK_ARG(a && a->M > 0 && a->N > 0 && a->K > 0 && a->nb_inner > 0 && a->nb_outer > 0);
  if (!operand_ok(a->A) || !operand_ok(a->B)) return DML_EINVAL;
  const int splits = a->splits > 1 ? a->splits : 1;
  const int softmax = a->softmax;
  DML_CHECK_ARG(softmax >= 0 && softmax <= 2);
  DML_CHECK_ARG(a->c || a->pair || a->half_out);
  if (splits > 1) {
    // split-K reduces alpha * partial products into a C the CALLER has zeroed: no other epilogue stage applies
    if (!a->c || a->pair || a->half_out || a->bias || a->resid || a->accumulate || a->relu || a->use_diag || softmax || a->absmax)
      return DML_EINVAL;
  }
  if (a->accumulate && !a->c) return DML_EINVAL;
  if (softmax == 2 && !a->aux) return DML_EINVAL;
  if (softmax && a->N > 256) return DML_EUNSUPPORTED;
  if (softmax && (a->bias || a->resid || a->accumulate || a->relu || a->use_diag || a->ncol_split > 0 || a->absmax)) return DML_EINVAL;
  if (a->pair && ((a->ldp % 8) || (a->p_plane % 8) || (((uintptr_t)a->pair) & 15))) return DML_EINVAL;
  int BN = softmax ? (a->N > 128 ? 256 : (a->N > 64 ? 128 : 64)) : (a->N > 64 ? 128 : 64);
  // a small problem (the 8 x 256^3 products of the pseudo-inverse: 32 tiles of 128 x 128) is latency-bound per CTA: half-width
  // tiles put it on twice the SMs with half the operand bytes, MMAs and epilogue columns each
  if (narrow_small && !softmax && BN == 128 &&
      2LL * cdiv(a->M, kBM) * cdiv(a->N, 128) * a->nb_inner * a->nb_outer * splits <= (long long)sm_count_cached())
    BN = 64;
  int rc;
  if ((rc = make_map5(ma, a->A, a->K, a->nb_inner, a->nb_outer, kBM)) || (rc = make_map5(mb, a->B, a->K, a->nb_inner, a->nb_outer, BN)))
    return rc;
  p = Params{};
  p.M = a->M; p.N = a->N; p.K = a->K; p.nb_inner = a->nb_inner; p.splits = splits;
  p.a_layout = a->A.layout; p.b_layout = a->B.layout;
  p.a_planes = a->A.plane_stride ? 2 : 1; p.b_planes = a->B.plane_stride ? 2 : 1;
  p.a_row_off = a->A.row_offset; p.a_k_off = a->A.k_offset; p.b_row_off = a->B.row_offset; p.b_k_off = a->B.k_offset;
  p.a_bi = a->A.bs_inner != 0; p.a_bo = a->A.bs_outer != 0; p.b_bi = a->B.bs_inner != 0; p.b_bo = a->B.bs_outer != 0;
  p.alpha = a->alpha; p.alpha2 = a->ncol_split > 0 ? a->alpha2 : 1.0f; p.ncol_split = a->ncol_split > 0 ? a->ncol_split : 0;
  p.alpha_dev = a->alpha_dev;
  p.bias = a->bias; p.bias_bi = a->bias_bs_inner; p.bias_bo = a->bias_bs_outer;
  p.relu = a->relu; p.use_diag = a->use_diag; p.diag = a->diag;
  p.resid = a->resid; p.ldr = a->ldr; p.r_bi = a->r_bs_inner; p.r_bo = a->r_bs_outer; p.resid_scale = a->resid_scale;
  p.accumulate = a->accumulate;
  p.c = a->c; p.ldc = a->ldc; p.c_bi = a->c_bs_inner; p.c_bo = a->c_bs_outer;
  p.pair = (bf16*)a->pair; p.ldp = a->ldp; p.p_bi = a->p_bs_inner; p.p_bo = a->p_bs_outer; p.p_plane = a->p_plane;
  p.half_out = (h16*)a->half_out; p.ldh = a->ldh; p.h_bi = a->h_bs_inner; p.h_bo = a->h_bs_outer; p.half_scale_dev = a->half_scale_dev;
  p.absmax = (uint32_t*)a->absmax;
  p.softmax = softmax;
  {
    auto al16 = [](const void* q) { return (((uintptr_t)q) & 15) == 0; };
    bool ok = true;
    if (a->c) ok = ok && al16(a->c) && (a->ldc % 4) == 0 && (a->c_bs_inner % 4) == 0 && (a->c_bs_outer % 4) == 0;
    if (a->resid) ok = ok && al16(a->resid) && (a->ldr % 4) == 0 && (a->r_bs_inner % 4) == 0 && (a->r_bs_outer % 4) == 0;
    if (a->bias) ok = ok && al16(a->bias) && (a->bias_bs_inner % 4) == 0 && (a->bias_bs_outer % 4) == 0;
    if (a->pair) ok = ok && (a->p_bs_inner % 8) == 0 && (a->p_bs_outer % 8) == 0;
    if (a->half_out) ok = ok && al16(a->half_out) && (a->ldh % 8) == 0 && (a->h_bs_inner % 8) == 0 && (a->h_bs_outer % 8) == 0;
    if (a->aux) ok = ok && al16(a->aux) && (a->ldx % 8) == 0 && (a->x_plane % 8) == 0 && (a->x_bs_inner % 8) == 0 && (a->x_bs_outer % 8) == 0;
    p.vec_ok = ok ? 1 : 0;
  }
  p.aux = (const bf16*)a->aux; p.ldx = a->ldx; p.x_bi = a->x_bs_inner; p.x_bo = a->x_bs_outer; p.x_plane = a->x_plane;
  const long long nz = (long long)a->nb_inner * a->nb_outer * splits;
  p.tiles_m = cdiv(a->M, kBM);
  p.tiles_n = cdiv(a->N, BN);
  const long long total = (long long)p.tiles_m * p.tiles_n * nz;
  if (total > 0x7fffffffLL) return DML_EUNSUPPORTED;
  p.total_tiles = (int)total;
  *bn_out = BN;
  return DML_OK;
}

static int device_sm_count() {
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)
    nsm = 148;
  return nsm;
}

int dml_pgemm(const dml_pgemm_args* a, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  using namespace dml::tc::pg;
  CUtensorMap ma, mb;
  Params p;
  int BN = 0;
  static const bool narrow = []() { const char* v = getenv("DML_B200_PGEMM_NARROW"); return !(v && v[0] == '0'); }();
  int rc = prepare_problem(a, p, &ma, &mb, &BN, narrow);
  if (rc) return rc;
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)min(device_sm_count(), p.total_tiles));
  cfg.blockDim = dim3(kThreads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see the kernel: griddepcontrol.wait guards every global access
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = []() { const char* v = getenv("DML_B200_PDL"); return !(v && v[0] == '0'); }();
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (BN == 64) {
    if ((e = cudaFuncSetAttribute(pgemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64>::kSmemBytes)) != cudaSuccess) return (int)e;
    cfg.dynamicSmemBytes = Cfg<64>::kSmemBytes;
    e = cudaLaunchKernelEx(&cfg, pgemm_kernel<64>, ma, mb, p);
  } else if (BN == 128) {
    if ((e = cudaFuncSetAttribute(pgemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::kSmemBytes)) != cudaSuccess) return (int)e;
    cfg.dynamicSmemBytes = Cfg<128>::kSmemBytes;
    e = cudaLaunchKernelEx(&cfg, pgemm_kernel<128>, ma, mb, p);
  } else {
    if ((e = cudaFuncSetAttribute(pgemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<256>::kSmemBytes)) != cudaSuccess) return (int)e;
    cfg.dynamicSmemBytes = Cfg<256>::kSmemBytes;
    e = cudaLaunchKernelEx(&cfg, pgemm_kernel<256>, ma, mb, p);
  }
  return e == cudaSuccess ? DML_OK : (int)e;
}

int dml_pgemm_chain_max(void) { return dml::tc::pg::kMaxChain; }

/* count <= dml_pgemm_chain_max() DEPENDENT problems in one cooperative launch: problem i + 1 may read what problem i wrote
 * (grid barrier between problems).  Every problem must need the 128-wide tile (64 < N, no fused softmax over more than 128
 * columns), have at most one tile per SM (tiles_m * tiles_n * batch <= SM count) and no split-K.                         */
int dml_pgemm_chain(const dml_pgemm_args* args, int count, void* stream) {
  using namespace dml;
  using namespace dml::tc;
  using namespace dml::tc::pg;
  DML_CHECK_ARG(args && count > 0 && count <= kMaxChain);
  static thread_local ChainParams cp;      // 14 KB: assembled here, copied into the launch (a kernel parameter) before returning
  cp.count = count;
  const int nsm = device_sm_count();
  int grid = 1;
  for (int i = 0; i < count; ++i) {
    int BN = 0;
    int rc = prepare_problem(&args[i], cp.p[i], &cp.maps[2 * i], &cp.maps[2 * i + 1], &BN, false);
    if (rc) return rc;
    if (BN != 128 || cp.p[i].splits != 1 || cp.p[i].total_tiles > nsm) return DML_EUNSUPPORTED;
    grid = max(grid, cp.p[i].total_tiles);
  }
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(pgemm_chain_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::kSmemBytes)) != cudaSuccess)
    return (int)e;
  void* kargs[1] = {(void*)&cp};
  e = cudaLaunchCooperativeKernel((void*)pgemm_chain_kernel<128>, dim3(grid), dim3(kThreads), kargs, Cfg<128>::kSmemBytes, (cudaStream_t)stream);
  return e == cudaSuccess ? DML_OK : (int)e;
}

int dml_pair_from_f32(const float* x, long long rows, int cols, int ld, float mult, void* pair, int ldp, long long plane_stride,
                      void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && pair && rows > 0 && cols > 0 && ld >= cols && ldp >= cols && (ldp % 4) == 0 && (plane_stride % 4) == 0);
  DML_CHECK_ARG((((uintptr_t)pair) & 7) == 0);
  const long long total = rows * ((cols + 3) / 4);
  const int blocks = (int)min((total + 255) / 256, (long long)148 * 16);
  tc::pg::pair_from_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, mult, (bf16*)pair, ldp, plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_relu_mask_pair(const float* g, float* gm, const float* act, long long rows, int cols, int ldg, int lda, void* pair, int ldp,
                       long long plane_stride, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(g && gm && ((((uintptr_t)gm) & 15) == 0) && act && pair && rows > 0 && cols > 0 && (cols % 4) == 0 && (ldg % 4) == 0 && (lda % 4) == 0 && (ldp % 4) == 0);
  DML_CHECK_ARG(((((uintptr_t)g) | ((uintptr_t)act)) & 15) == 0 && (((uintptr_t)pair) & 7) == 0 && (plane_stride % 4) == 0);
  const long long total = rows * (cols / 4);
  const int blocks = (int)min((total + 255) / 256, (long long)148 * 16);
  tc::pg::relu_mask_pair_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, gm, act, rows, cols, ldg, lda, (bf16*)pair, ldp, plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_loss_scale_from_amax(const void* amax_bits, float* scale2, void* stream) {
  DML_CHECK_ARG(amax_bits && scale2);
  dml::tc::pg::loss_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned*)amax_bits, scale2);
  DML_RETURN_LAUNCH();
}

int dml_scale_to_half(const float* x, const float* scale_dev, long long n, void* out, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && scale_dev && out && n > 0 && (n % 8) == 0 && ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0);
  const long long n8 = n / 8;
  const int blocks = (int)min((n8 + 255) / 256, (long long)148 * 16);
  tc::pg::scale_to_half_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, scale_dev, n8, (h16*)out);
  DML_RETURN_LAUNCH();
}

int dml_colsum(const float* x, long long rows, int cols, int ld, float* out, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, st);
  if (e != cudaSuccess) return (int)e;
  const int gy = (int)max(1LL, min((long long)148 * 4 / cdiv(cols, 32), (rows + 63) / 64));
  tc::pg::colsum_kernel<<<dim3(cdiv(cols, 32), gy), 256, 0, st>>>(x, rows, cols, ld, out);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
