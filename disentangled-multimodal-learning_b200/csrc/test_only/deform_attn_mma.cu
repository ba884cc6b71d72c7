// Fused deformable cross-attention with on-the-fly continuous-position bias (v1: warp-level mma.sync).
//
// Reference maths (DeformableAttention1D.py:203-231 + CPB :84-102):
//     S[h,i,j] = scale * q[h,i,:].k[h,j,:] + bias_{h % O}( sign(p) log(|p|+1) ),  p = seq_i - g[(b, h / O), j]
//     P = softmax_j(S);  out[h,i,:] = sum_j P[h,i,j] v[h,j,:]
// sim/bias/attn ([8, n, n_kv] each, 2-34 GB at n = 16k in the reference) are never materialised: the
// forward keeps a running (max, sum) per row, the backward recomputes P from the saved log-sum-exp.
// The bias is evaluated through the exact piecewise-linear table of cpb_table.cuh.
//
// Three kernels: forward (query-stationary), backward dK/dV/dg/segment-sums (key-stationary, S^T
// orientation so that each thread owns fixed keys j and sees monotonically increasing i -> the
// per-segment sums are run-length merged in registers), backward dQ (query-stationary).
// Tensor-core path here is mma.sync.m16n8k16 (fp16 in, fp32 accumulate); the tcgen05/TMEM version
// replaces these kernels behind the same C entry points.
#include <math.h>

#include "../../../include/dml_b200.h"
#include "../../../include/dml_b200_test.h"
#include "../common.cuh"
#include "../cpb_table.cuh"

namespace dml {

constexpr int kD = 64;      // head dim
constexpr int kBM = 64;     // rows per CTA (4 warps x 16)
constexpr int kBN = 64;     // streamed tile
constexpr int kTile = 64 * 64;

struct AttnParams {
  const h16* q; const h16* k; const h16* v;     // [B,n,ldq] / [B,n_kv,ldk] / [B,n_kv,ldv], head h at column h*64
  const float* g;                                   // [(B G), n_kv] normalised sampling positions
  const uint32_t* table;                            // cpb table (device)
  float* o; float* lse;                             // fp32 [B,n,ldo], [B,H,n] (log2 domain)
  const h16* d_o; const float* dsum;               // backward: dO [B,n,ldo], D [B,H,n]
  float* dq; float* dk; float* dv; float* dg;       // fp32 [B,n,H*64], [B,n_kv,H*64] x2, [(B G), n_kv]
  float* segsum;                                    // [kCpbSegMax][4]
  const float* dscale;                              // device [2] = (s, 1/s): d_o (hence D, dS) arrive multiplied by s
  int B, H, n, n_kv, ldq, ldk, ldv, ldo, nout;
  float scale;
};

__device__ __forceinline__ float seq_pos(int i, int n) {
  // normalize_grid(arange(n)) in fp32, same operation order as the reference (:45-48)
  return (2.0f * (float)i) / (float)max(n - 1, 1) - 1.0f;
}

// 64-row x 64-col bf16 tile -> swizzled smem; rows >= nrows zero-filled.  128 threads.
__device__ __forceinline__ void load_tile64(h16* dst, const h16* src, int ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int id = tid + 128 * k, r = id >> 3, ch = id & 7;
    const bool ok = (row0 + r) < nrows;
    const h16* s = src + (size_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(smem_u32(dst + swz64(r, ch)), s, ok);
  }
}
// 32-row variant
__device__ __forceinline__ void load_tile32(h16* dst, const h16* src, int ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int id = tid + 128 * k, r = id >> 3, ch = id & 7;
    const bool ok = (row0 + r) < nrows;
    const h16* s = src + (size_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(smem_u32(dst + swz64(r, ch)), s, ok);
  }
}

// A fragment (16 rows x 16 k) of a swizzled [rows][64] tile: rows r0.., k-chunk pair kc (k = 16*kc..)
__device__ __forceinline__ void lda(uint32_t (&a)[4], const h16* tile, int r0, int kc, int lane) {
  const int r = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = kc * 2 + (lane >> 4);
  ldmatrix_x4(a, smem_u32(tile + swz64(r, ch)));
}
// B fragments for two n-tiles (16 n) x 16 k from a tile stored [n][k] (k contiguous): n0.., k-chunk pair kc
__device__ __forceinline__ void ldb_nk(uint32_t (&b)[4], const h16* tile, int n0, int kc, int lane) {
  const int r = n0 + (lane & 7) + (lane >> 4) * 8, ch = kc * 2 + ((lane >> 3) & 1);
  ldmatrix_x4(b, smem_u32(tile + swz64(r, ch)));
}
// B fragments for two n-tiles (16 n) x 16 k from a tile stored [k][n] (n contiguous): k0.., n-chunk pair nc
__device__ __forceinline__ void ldb_kn(uint32_t (&b)[4], const h16* tile, int k0, int nc, int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = nc * 2 + (lane >> 4);
  ldmatrix_x4_trans(b, smem_u32(tile + swz64(r, ch)));
}

// ================================================================================================
// forward
// ================================================================================================
constexpr int kFwdSmem = 4 * kTile * 2 + 2 * kBN * 4 + kCpbSmemFwdBytes;

// scale + bias (+ column mask) for one 16x64 score fragment, log2 domain; returns the row maxima
template <bool kMask>
__device__ __forceinline__ void fwd_bias(float (&s)[8][4], const CpbView& tb, const float* gt, float s_lo, float s_hi,
                                         float sc2, int lane, int jrem, float& mx_lo, float& mx_hi) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int jj = nt * 8 + (lane & 3) * 2 + e;
      const float gj = gt[jj];
      float a, c;
      int dummy;
      const float x0 = cpb_x(s_lo - gj);
      cpb_lookup<false>(tb, x0, a, c, dummy);
      float v0 = fmaf(s[nt][e], sc2, fmaf(a, x0, c));
      const float x1 = cpb_x(s_hi - gj);
      cpb_lookup<false>(tb, x1, a, c, dummy);
      float v1 = fmaf(s[nt][2 + e], sc2, fmaf(a, x1, c));
      if (kMask && jj >= jrem) { v0 = -INFINITY; v1 = -INFINITY; }
      s[nt][e] = v0;
      s[nt][2 + e] = v1;
      mx_lo = fmaxf(mx_lo, v0);
      mx_hi = fmaxf(mx_hi, v1);
    }
  }
}

__global__ void __launch_bounds__(128, 3) deform_attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  h16* Ks = reinterpret_cast<h16*>(smem);                 // 2 stages
  h16* Vs = Ks + 2 * kTile;                                // 2 stages (stage 1 holds the Q tile during the prologue)
  float* gs = reinterpret_cast<float*>(Vs + 2 * kTile);
  uint8_t* tabs = reinterpret_cast<uint8_t*>(gs + 2 * kBN);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const h16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const h16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const h16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
  const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
  const int ntiles = cdiv(p.n_kv, kBN);

  load_tile64(Vs + kTile, qb, p.ldq, i0, p.n, tid);
  load_tile64(Ks, kb, p.ldk, 0, p.n_kv, tid);
  load_tile64(Vs, vb, p.ldv, 0, p.n_kv, tid);
  if (tid < kBN) cp_async4(smem_u32(gs + tid), gb + min(tid, p.n_kv - 1), tid < p.n_kv);
  cp_async_commit();
  const CpbView tb = cpb_stage(tabs, p.table, oidx, false, tid, 128);
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4];
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) lda(qf[kc], Vs + kTile, warp * 16, kc, lane);

  const int r_lo = warp * 16 + (lane >> 2);
  const float s_lo = seq_pos(i0 + r_lo, p.n), s_hi = seq_pos(i0 + r_lo + 8, p.n);
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const float sc2 = p.scale * kLog2e;

  for (int jt = 0; jt < ntiles; ++jt) {
    const int buf = jt & 1;
    cp_async_wait<0>();
    __syncthreads();                       // tile jt landed; every warp is done with tile jt-1 (and with the Q tile)
    if (jt + 1 < ntiles) {
      const int j1 = (jt + 1) * kBN;
      load_tile64(Ks + (buf ^ 1) * kTile, kb, p.ldk, j1, p.n_kv, tid);
      load_tile64(Vs + (buf ^ 1) * kTile, vb, p.ldv, j1, p.n_kv, tid);
      if (tid < kBN) cp_async4(smem_u32(gs + (buf ^ 1) * kBN + tid), gb + min(j1 + tid, p.n_kv - 1), j1 + tid < p.n_kv);
      cp_async_commit();
    }
    const h16* Kt = Ks + buf * kTile;
    const h16* Vt = Vs + buf * kTile;
    const float* gt = gs + buf * kBN;

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_nk(bb, Kt, np * 16, kc, lane);
        mma_f16_16816(s[2 * np], qf[kc], bb[0], bb[1]);
        mma_f16_16816(s[2 * np + 1], qf[kc], bb[2], bb[3]);
      }
    }
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
    const int jrem = p.n_kv - jt * kBN;
    if (jrem >= kBN) fwd_bias<false>(s, tb, gt, s_lo, s_hi, sc2, lane, jrem, mx_lo, mx_hi);
    else fwd_bias<true>(s, tb, gt, s_lo, s_hi, sc2, lane, jrem, mx_lo, mx_hi);
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float al_lo = exp2f(m_lo - mn_lo), al_hi = exp2f(m_hi - mn_hi);
    m_lo = mn_lo;
    m_hi = mn_hi;
    float rs_lo = 0.f, rs_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn_lo); s[nt][1] = exp2f(s[nt][1] - mn_lo);
      s[nt][2] = exp2f(s[nt][2] - mn_hi); s[nt][3] = exp2f(s[nt][3] - mn_hi);
      rs_lo += s[nt][0] + s[nt][1];
      rs_hi += s[nt][2] + s[nt][3];
    }
    l_lo = l_lo * al_lo + rs_lo;
    l_hi = l_hi * al_hi + rs_hi;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= al_lo; o[nt][1] *= al_lo; o[nt][2] *= al_hi; o[nt][3] *= al_hi;
    }
    // O += P V with P = P_hi + P_lo (two fp16 terms, 22 significant bits): O, and with it the backward's
    // D = dO.O = sum_j P_ij dP_ij, carry no P-rounding error, so sum_j dS_ij = 0 holds to fp32 accuracy and the
    // cancellation-dominated bias-MLP gradients are not polluted by a row-coherent error term.
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4], al[4];
      split_f16(s[2 * kk][0], s[2 * kk][1], a[0], al[0]);
      split_f16(s[2 * kk][2], s[2 * kk][3], a[1], al[1]);
      split_f16(s[2 * kk + 1][0], s[2 * kk + 1][1], a[2], al[2]);
      split_f16(s[2 * kk + 1][2], s[2 * kk + 1][3], a[3], al[3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bb[4];
        ldb_kn(bb, Vt, kk * 16, dp, lane);
        mma_f16_16816(o[2 * dp], a, bb[0], bb[1]);
        mma_f16_16816(o[2 * dp + 1], a, bb[2], bb[3]);
        mma_f16_16816(o[2 * dp], al, bb[0], bb[1]);
        mma_f16_16816(o[2 * dp + 1], al, bb[2], bb[3]);
      }
    }
  }
  // finalize
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
  const int gi_lo = i0 + r_lo, gi_hi = gi_lo + 8;
  float* ob = p.o + (size_t)b * p.n * p.ldo + h * kD + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (gi_lo < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_lo * p.ldo + nt * 8) = make_float2(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    if (gi_hi < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_hi * p.ldo + nt * 8) = make_float2(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
  if ((lane & 3) == 0) {
    float* lb = p.lse + ((size_t)b * p.H + h) * p.n;
    if (gi_lo < p.n) lb[gi_lo] = m_lo + log2f(l_lo);
    if (gi_hi < p.n) lb[gi_hi] = m_hi + log2f(l_hi);
  }
}

// ================================================================================================
// backward prep: D[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]   (O in fp32 as written by the forward, dO the bf16
// tensor the MMAs consume -> the softmax-backward identity sum_j dS_ij = 0 holds to fp32 accuracy)
// ================================================================================================
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const float* __restrict__ o, const h16* __restrict__ d_o,
                                                            int B, int n, int H, int ld, float* __restrict__ dsum) {
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

// ================================================================================================
// backward dQ (query-stationary): dQ[i,:] = sum_j dS[i,j] K[j,:]   (unscaled; the caller applies `scale`)
// ================================================================================================
constexpr int kDqSmem = 4 * kTile * 2 + 2 * kBN * 4 + kCpbSmemFwdBytes;

template <bool kMask>
__device__ __forceinline__ void dq_ds(float (&s)[8][4], const float (&dp)[8][4], const CpbView& tb, const float* gt,
                                      float s_lo, float s_hi, float sc2, float lse_lo, float lse_hi, float d_lo,
                                      float d_hi, int lane, int jrem) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int jj = nt * 8 + (lane & 3) * 2 + e;
      const float gj = gt[jj];
      float a, c;
      int dummy;
      const float x0 = cpb_x(s_lo - gj);
      cpb_lookup<false>(tb, x0, a, c, dummy);
      float p0 = exp2f(fmaf(s[nt][e], sc2, fmaf(a, x0, c)) - lse_lo);
      const float x1 = cpb_x(s_hi - gj);
      cpb_lookup<false>(tb, x1, a, c, dummy);
      float p1 = exp2f(fmaf(s[nt][2 + e], sc2, fmaf(a, x1, c)) - lse_hi);
      if (kMask && jj >= jrem) { p0 = 0.f; p1 = 0.f; }
      s[nt][e] = p0 * (dp[nt][e] - d_lo);
      s[nt][2 + e] = p1 * (dp[nt][2 + e] - d_hi);
    }
  }
}

__global__ void __launch_bounds__(128, 3) deform_attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  h16* Ks = reinterpret_cast<h16*>(smem);   // 2 stages (stage 1 holds the Q tile during the prologue)
  h16* Vs = Ks + 2 * kTile;                  // 2 stages (stage 1 holds the dO tile during the prologue)
  float* gs = reinterpret_cast<float*>(Vs + 2 * kTile);
  uint8_t* tabs = reinterpret_cast<uint8_t*>(gs + 2 * kBN);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const h16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const h16* dob = p.d_o + (size_t)b * p.n * p.ldo + h * kD;
  const h16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const h16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
  const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
  const int ntiles = cdiv(p.n_kv, kBN);

  load_tile64(Ks + kTile, qb, p.ldq, i0, p.n, tid);
  load_tile64(Vs + kTile, dob, p.ldo, i0, p.n, tid);
  load_tile64(Ks, kb, p.ldk, 0, p.n_kv, tid);
  load_tile64(Vs, vb, p.ldv, 0, p.n_kv, tid);
  if (tid < kBN) cp_async4(smem_u32(gs + tid), gb + min(tid, p.n_kv - 1), tid < p.n_kv);
  cp_async_commit();
  const CpbView tb = cpb_stage(tabs, p.table, oidx, false, tid, 128);
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    lda(qf[kc], Ks + kTile, warp * 16, kc, lane);
    lda(dof[kc], Vs + kTile, warp * 16, kc, lane);
  }

  const int r_lo = warp * 16 + (lane >> 2);
  const int gi_lo = i0 + r_lo, gi_hi = gi_lo + 8;
  const float s_lo = seq_pos(gi_lo, p.n), s_hi = seq_pos(gi_hi, p.n);
  const float* lb = p.lse + ((size_t)b * p.H + h) * p.n;
  const float* db = p.dsum + ((size_t)b * p.H + h) * p.n;
  const float lse_lo = gi_lo < p.n ? lb[gi_lo] : 0.f, lse_hi = gi_hi < p.n ? lb[gi_hi] : 0.f;
  const float d_lo = gi_lo < p.n ? db[gi_lo] : 0.f, d_hi = gi_hi < p.n ? db[gi_hi] : 0.f;
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const float sc2 = p.scale * kLog2e;

  for (int jt = 0; jt < ntiles; ++jt) {
    const int buf = jt & 1;
    cp_async_wait<0>();
    __syncthreads();
    if (jt + 1 < ntiles) {
      const int j1 = (jt + 1) * kBN;
      load_tile64(Ks + (buf ^ 1) * kTile, kb, p.ldk, j1, p.n_kv, tid);
      load_tile64(Vs + (buf ^ 1) * kTile, vb, p.ldv, j1, p.n_kv, tid);
      if (tid < kBN) cp_async4(smem_u32(gs + (buf ^ 1) * kBN + tid), gb + min(j1 + tid, p.n_kv - 1), j1 + tid < p.n_kv);
      cp_async_commit();
    }
    const h16* Kt = Ks + buf * kTile;
    const h16* Vt = Vs + buf * kTile;
    const float* gt = gs + buf * kBN;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_nk(bb, Kt, np * 16, kc, lane);
        mma_f16_16816(s[2 * np], qf[kc], bb[0], bb[1]);
        mma_f16_16816(s[2 * np + 1], qf[kc], bb[2], bb[3]);
        ldb_nk(bb, Vt, np * 16, kc, lane);
        mma_f16_16816(dp[2 * np], dof[kc], bb[0], bb[1]);
        mma_f16_16816(dp[2 * np + 1], dof[kc], bb[2], bb[3]);
      }
    }
    const int jrem = p.n_kv - jt * kBN;
    if (jrem >= kBN) dq_ds<false>(s, dp, tb, gt, s_lo, s_hi, sc2, lse_lo, lse_hi, d_lo, d_hi, lane, jrem);
    else dq_ds<true>(s, dp, tb, gt, s_lo, s_hi, sc2, lse_lo, lse_hi, d_lo, d_hi, lane, jrem);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_f16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_f16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_f16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_f16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bb[4];
        ldb_kn(bb, Kt, kk * 16, dpi, lane);
        mma_f16_16816(dq[2 * dpi], a, bb[0], bb[1]);
        mma_f16_16816(dq[2 * dpi + 1], a, bb[2], bb[3]);
      }
    }
  }
  float* ob = p.dq + (size_t)b * p.n * (p.H * kD) + h * kD + (lane & 3) * 2;
  const float inv_s = __ldg(p.dscale + 1);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (gi_lo < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_lo * (p.H * kD) + nt * 8) = make_float2(dq[nt][0] * inv_s, dq[nt][1] * inv_s);
    if (gi_hi < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_hi * (p.H * kD) + nt * 8) = make_float2(dq[nt][2] * inv_s, dq[nt][3] * inv_s);
  }
}

// ================================================================================================
// backward dK / dV / dg / CPB segment sums (key-stationary, transposed orientation: rows = keys j)
// ================================================================================================
constexpr int kQT = 32;            // query rows per streamed tile
constexpr int kQTile = kQT * 64;
constexpr int kStages = 3;
constexpr int kDkvSmem = kStages * 2 * kQTile * 2 + kStages * 2 * kQT * 4 + kCpbSmemBwdBytes + kCpbSegMax * 2 * 4;

__global__ void __launch_bounds__(128, 2) deform_attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  h16* Qs = reinterpret_cast<h16*>(smem);                // kStages x [32][64]; stages 0/1 hold the K tile in the prologue
  h16* dOs = Qs + kStages * kQTile;                       // kStages x [32][64]; stages 0/1 hold the V tile in the prologue
  float* ls = reinterpret_cast<float*>(dOs + kStages * kQTile);  // kStages x (lse[32], D[32])
  uint8_t* tabs = reinterpret_cast<uint8_t*>(ls + kStages * 2 * kQT);
  float* ssum = reinterpret_cast<float*>(tabs + kCpbSmemBwdBytes);   // [kCpbSegMax][2]  (A, Bx) for this head's output

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const h16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const h16* dob = p.d_o + (size_t)b * p.n * p.ldo + h * kD;
  const h16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const h16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
  const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
  const float* lb = p.lse + ((size_t)b * p.H + h) * p.n;
  const float* db = p.dsum + ((size_t)b * p.H + h) * p.n;
  const int ntiles = cdiv(p.n, kQT);

  auto load_stage = [&](int st, int it) {
    const int r0 = it * kQT;
    load_tile32(Qs + st * kQTile, qb, p.ldq, r0, p.n, tid);
    load_tile32(dOs + st * kQTile, dob, p.ldo, r0, p.n, tid);
    if (tid < kQT) cp_async4(smem_u32(ls + st * 2 * kQT + tid), lb + min(r0 + tid, p.n - 1), r0 + tid < p.n);
    else if (tid < 2 * kQT) cp_async4(smem_u32(ls + st * 2 * kQT + tid), db + min(r0 + tid - kQT, p.n - 1), r0 + tid - kQT < p.n);
  };

  // prologue: the 64-row K and V tiles pass through the (still unused) stage buffers into register fragments
  load_tile64(Qs, kb, p.ldk, j0, p.n_kv, tid);      // 64x64 = stages 0+1 of Qs
  load_tile64(dOs, vb, p.ldv, j0, p.n_kv, tid);     // 64x64 = stages 0+1 of dOs
  cp_async_commit();
  const CpbView tb = cpb_stage(tabs, p.table, oidx, true, tid, 128);
  for (int i = tid; i < kCpbSegMax * 2; i += 128) ssum[i] = 0.f;
  cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    lda(kf[kc], Qs, warp * 16, kc, lane);
    lda(vf[kc], dOs, warp * 16, kc, lane);
  }
  __syncthreads();
  load_stage(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_stage(1, 1);
  cp_async_commit();

  const int r_lo = warp * 16 + (lane >> 2);
  const int gj_lo = j0 + r_lo, gj_hi = gj_lo + 8;
  const bool jv_lo = gj_lo < p.n_kv, jv_hi = gj_hi < p.n_kv;
  const float g_lo = gb[min(gj_lo, p.n_kv - 1)], g_hi = gb[min(gj_hi, p.n_kv - 1)];
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  float dg_lo = 0.f, dg_hi = 0.f;
  int run_seg[2] = {-1, -1};
  float run_a[2] = {0.f, 0.f}, run_b[2] = {0.f, 0.f};
  const float sc2 = p.scale * kLog2e;
  const float inv_den = 1.0f / (float)max(p.n - 1, 1);

  for (int it = 0; it < ntiles; ++it) {
    const int st = it % kStages;
    cp_async_wait<1>();
    __syncthreads();                 // tile `it` landed; every warp is done with tile it-1 -> its stage may be refilled
    if (it + 2 < ntiles) load_stage((it + 2) % kStages, it + 2);
    cp_async_commit();
    const h16* Qt = Qs + st * kQTile;
    const h16* dOt = dOs + st * kQTile;
    const float* lt = ls + st * 2 * kQT;

    // S^T = K Q^T  and  dP^T = V dO^T   ([16 j] x [32 i] per warp)
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
      for (int ip = 0; ip < 2; ++ip) {
        uint32_t bb[4];
        ldb_nk(bb, Qt, ip * 16, kc, lane);
        mma_f16_16816(s[2 * ip], kf[kc], bb[0], bb[1]);
        mma_f16_16816(s[2 * ip + 1], kf[kc], bb[2], bb[3]);
        ldb_nk(bb, dOt, ip * 16, kc, lane);
        mma_f16_16816(dp[2 * ip], vf[kc], bb[0], bb[1]);
        mma_f16_16816(dp[2 * ip + 1], vf[kc], bb[2], bb[3]);
      }
    }
    // P^T, dS^T, bias gradients
    const int ibase = it * kQT;
    float pT[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ii = nt * 8 + (lane & 3) * 2 + e;
        const int gi = ibase + ii;
        const bool iv = gi < p.n;
        const float si = (2.0f * (float)gi) * inv_den - 1.0f;
        const float lse_i = lt[ii], d_i = lt[kQT + ii];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float gj = r ? g_hi : g_lo;
          const bool valid = iv && (r ? jv_hi : jv_lo);
          const float x = cpb_x(si - gj);
          float a_s, c_s;
          int seg;
          cpb_lookup<true>(tb, x, a_s, c_s, seg);
          const float pr = valid ? exp2f(fmaf(s[nt][2 * r + e], sc2, fmaf(a_s, x, c_s)) - lse_i) : 0.f;
          const float ds = pr * (dp[nt][2 * r + e] - d_i);
          pT[nt][2 * r + e] = pr;
          s[nt][2 * r + e] = ds;
          // d bias / d g_j = -a_s / (|p| + 1),  |p| + 1 = 2^|x|
          const float dgc = -ds * a_s * exp2f(-fabsf(x));
          if (r) dg_hi += dgc; else dg_lo += dgc;
          if (seg != run_seg[r]) {
            if (run_seg[r] >= 0 && (run_a[r] != 0.f || run_b[r] != 0.f)) {
              atomicAdd(ssum + 2 * run_seg[r], run_a[r]);
              atomicAdd(ssum + 2 * run_seg[r] + 1, run_b[r]);
            }
            run_seg[r] = seg;
            run_a[r] = 0.f;
            run_b[r] = 0.f;
          }
          run_a[r] += ds;
          run_b[r] = fmaf(ds, x, run_b[r]);
        }
      }
    }
    // dV += P^T dO ;  dK += dS^T Q      (k = i, two k-steps of 16)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t ap[4], as[4];
      ap[0] = pack_f16(pT[2 * kk][0], pT[2 * kk][1]);
      ap[1] = pack_f16(pT[2 * kk][2], pT[2 * kk][3]);
      ap[2] = pack_f16(pT[2 * kk + 1][0], pT[2 * kk + 1][1]);
      ap[3] = pack_f16(pT[2 * kk + 1][2], pT[2 * kk + 1][3]);
      as[0] = pack_f16(s[2 * kk][0], s[2 * kk][1]);
      as[1] = pack_f16(s[2 * kk][2], s[2 * kk][3]);
      as[2] = pack_f16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      as[3] = pack_f16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bb[4];
        ldb_kn(bb, dOt, kk * 16, dpi, lane);
        mma_f16_16816(dv[2 * dpi], ap, bb[0], bb[1]);
        mma_f16_16816(dv[2 * dpi + 1], ap, bb[2], bb[3]);
        ldb_kn(bb, Qt, kk * 16, dpi, lane);
        mma_f16_16816(dk[2 * dpi], as, bb[0], bb[1]);
        mma_f16_16816(dk[2 * dpi + 1], as, bb[2], bb[3]);
      }
    }
  }
  // flush the run-length accumulators, then the CTA's segment sums
#pragma unroll
  for (int r = 0; r < 2; ++r)
    if (run_seg[r] >= 0 && (run_a[r] != 0.f || run_b[r] != 0.f)) {
      atomicAdd(ssum + 2 * run_seg[r], run_a[r]);
      atomicAdd(ssum + 2 * run_seg[r] + 1, run_b[r]);
    }
  __syncthreads();
  const int nseg = (int)__ldg(p.table);
  const float inv_s = __ldg(p.dscale + 1);
  for (int i = tid; i < nseg * 2; i += 128) {
    const float v = ssum[i];
    if (v != 0.f) atomicAdd(p.segsum + (i >> 1) * 4 + oidx * 2 + (i & 1), v * inv_s);
  }
  // dg: reduce over the 4 lanes that share a key row, one atomic per (head, key)
  dg_lo += __shfl_xor_sync(0xffffffffu, dg_lo, 1);
  dg_lo += __shfl_xor_sync(0xffffffffu, dg_lo, 2);
  dg_hi += __shfl_xor_sync(0xffffffffu, dg_hi, 1);
  dg_hi += __shfl_xor_sync(0xffffffffu, dg_hi, 2);
  if ((lane & 3) == 0) {
    float* dgb = p.dg + (size_t)(b * G + grp) * p.n_kv;
    if (jv_lo) atomicAdd(dgb + gj_lo, dg_lo * inv_s);
    if (jv_hi) atomicAdd(dgb + gj_hi, dg_hi * inv_s);
  }
  const int ldg = p.H * kD;
  const float ksc = p.scale * inv_s;
  float* dkb = p.dk + (size_t)b * p.n_kv * ldg + h * kD + (lane & 3) * 2;
  float* dvb = p.dv + (size_t)b * p.n_kv * ldg + h * kD + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (jv_lo) {
      *reinterpret_cast<float2*>(dkb + (size_t)gj_lo * ldg + nt * 8) = make_float2(dk[nt][0] * ksc, dk[nt][1] * ksc);
      *reinterpret_cast<float2*>(dvb + (size_t)gj_lo * ldg + nt * 8) = make_float2(dv[nt][0] * inv_s, dv[nt][1] * inv_s);
    }
    if (jv_hi) {
      *reinterpret_cast<float2*>(dkb + (size_t)gj_hi * ldg + nt * 8) = make_float2(dk[nt][2] * ksc, dk[nt][3] * ksc);
      *reinterpret_cast<float2*>(dvb + (size_t)gj_hi * ldg + nt * 8) = make_float2(dv[nt][2] * inv_s, dv[nt][3] * inv_s);
    }
  }
}

static int check_common(const AttnParams& p) {
  if (p.B <= 0 || p.H <= 0 || p.n <= 0 || p.n_kv <= 0) return DML_EINVAL;
  if (p.nout < 1 || p.nout > kCpbOutMax || (p.H % p.nout) != 0) return DML_EUNSUPPORTED;
  if ((p.ldq % 8) || (p.ldk % 8) || (p.ldv % 8) || (p.ldo % 8)) return DML_EINVAL;
  if (p.ldq < p.H * kD || p.ldk < p.H * kD || p.ldv < p.H * kD || p.ldo < p.H * kD) return DML_EINVAL;
  return DML_OK;
}

}  // namespace dml

extern "C" {

int dml_deform_attn_fwd(const void* q, const void* k, const void* v, const float* g, const void* table, int B, int H,
                        int dim_head, int n, int n_kv, int ldq, int ldk, int ldv, int ldo, int heads_per_group,
                        float scale, void* out, float* lse, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(q && k && v && g && table && out && lse);
  if (dim_head != kD) return DML_EUNSUPPORTED;
  AttnParams p{};
  p.q = (const h16*)q; p.k = (const h16*)k; p.v = (const h16*)v; p.g = g; p.table = (const uint32_t*)table;
  p.o = (float*)out; p.lse = lse;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.nout = heads_per_group;
  p.scale = scale;
  int rc = check_common(p);
  if (rc) return rc;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(deform_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(cdiv(n, kBM), H, B);
  deform_attn_fwd_kernel<<<grid, 128, kFwdSmem, (cudaStream_t)stream>>>(p);
  DML_RETURN_LAUNCH();
}

int dml_deform_attn_bwd(const void* q, const void* k, const void* v, const float* g, const void* table,
                        const void* out, const void* d_out, const float* lse, int B, int H, int dim_head, int n,
                        int n_kv, int ldq, int ldk, int ldv, int ldo, int heads_per_group, float scale,
                        const float* dscale, float* dsum_ws, float* dq, float* dk, float* dv, float* dg,
                        float* segsum, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(q && k && v && g && table && out && d_out && lse && dscale && dsum_ws && dq && dk && dv && dg && segsum);
  if (dim_head != kD) return DML_EUNSUPPORTED;
  AttnParams p{};
  p.q = (const h16*)q; p.k = (const h16*)k; p.v = (const h16*)v; p.g = g; p.table = (const uint32_t*)table;
  p.o = (float*)const_cast<void*>(out); p.lse = const_cast<float*>(lse); p.d_o = (const h16*)d_out; p.dsum = dsum_ws;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dg = dg; p.segsum = segsum; p.dscale = dscale;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.nout = heads_per_group;
  p.scale = scale;
  int rc = check_common(p);
  if (rc) return rc;
  if (ldo != H * kD) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(deform_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(deform_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
    if (e != cudaSuccess) return (int)e;
  }
  const int G = H / heads_per_group;
  cudaError_t e = cudaMemsetAsync(dg, 0, sizeof(float) * (size_t)B * G * n_kv, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(segsum, 0, sizeof(float) * 4 * kCpbSegMax, st);
  if (e != cudaSuccess) return (int)e;
  const int rows = B * n;
  attn_bwd_prep_kernel<<<min(cdiv(rows, 8), 148 * 8), 256, 0, st>>>((const float*)out, (const h16*)d_out, B, n, H, ldo, dsum_ws);
  deform_attn_bwd_dkv_kernel<<<dim3(cdiv(n_kv, kBN), H, B), 128, kDkvSmem, st>>>(p);
  deform_attn_bwd_dq_kernel<<<dim3(cdiv(n, kBM), H, B), 128, kDqSmem, st>>>(p);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
