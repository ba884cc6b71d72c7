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
// Tensor-core path here is mma.sync.m16n8k16 (bf16 in, fp32 accumulate); the tcgen05/TMEM version
// replaces these kernels behind the same C entry points.
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"
#include "cpb_table.cuh"

namespace dml {

constexpr int kD = 64;      // head dim
constexpr int kBM = 64;     // rows per CTA (4 warps x 16)
constexpr int kBN = 64;     // streamed tile
constexpr int kTile = 64 * 64;

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v;     // [B,n,ldq] / [B,n_kv,ldk] / [B,n_kv,ldv], head h at column h*64
  const float* g;                                   // [(B G), n_kv] normalised sampling positions
  const uint32_t* table;                            // cpb table (device)
  bf16* o; float* lse;                              // [B,n,ldo], [B,H,n] (log2 domain)
  const bf16* d_o; const float* dsum;               // backward: dO [B,n,ldo], D [B,H,n]
  float* dq; float* dk; float* dv; float* dg;       // fp32 [B,n,H*64], [B,n_kv,H*64] x2, [(B G), n_kv]
  float* segsum;                                    // [kCpbSegMax][4]
  int B, H, n, n_kv, ldq, ldk, ldv, ldo, nout;
  float scale;
};

__device__ __forceinline__ float seq_pos(int i, int n) {
  // normalize_grid(arange(n)) in fp32, same operation order as the reference (:45-48)
  return (2.0f * (float)i) / (float)max(n - 1, 1) - 1.0f;
}

// 64-row x 64-col bf16 tile -> swizzled smem; rows >= nrows zero-filled.  128 threads.
__device__ __forceinline__ void load_tile64(bf16* dst, const bf16* src, int ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int id = tid + 128 * k, r = id >> 3, ch = id & 7;
    const bool ok = (row0 + r) < nrows;
    const bf16* s = src + (size_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(smem_u32(dst + swz64(r, ch)), s, ok);
  }
}
// 32-row variant
__device__ __forceinline__ void load_tile32(bf16* dst, const bf16* src, int ld, int row0, int nrows, int tid) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int id = tid + 128 * k, r = id >> 3, ch = id & 7;
    const bool ok = (row0 + r) < nrows;
    const bf16* s = src + (size_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(smem_u32(dst + swz64(r, ch)), s, ok);
  }
}

// A fragment (16 rows x 16 k) of a swizzled [rows][64] tile: rows r0.., k-chunk pair kc (k = 16*kc..)
__device__ __forceinline__ void lda(uint32_t (&a)[4], const bf16* tile, int r0, int kc, int lane) {
  const int r = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = kc * 2 + (lane >> 4);
  ldmatrix_x4(a, smem_u32(tile + swz64(r, ch)));
}
// B fragments for two n-tiles (16 n) x 16 k from a tile stored [n][k] (k contiguous): n0.., k-chunk pair kc
__device__ __forceinline__ void ldb_nk(uint32_t (&b)[4], const bf16* tile, int n0, int kc, int lane) {
  const int r = n0 + (lane & 7) + (lane >> 4) * 8, ch = kc * 2 + ((lane >> 3) & 1);
  ldmatrix_x4(b, smem_u32(tile + swz64(r, ch)));
}
// B fragments for two n-tiles (16 n) x 16 k from a tile stored [k][n] (n contiguous): k0.., n-chunk pair nc
__device__ __forceinline__ void ldb_kn(uint32_t (&b)[4], const bf16* tile, int k0, int nc, int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, ch = nc * 2 + (lane >> 4);
  ldmatrix_x4_trans(b, smem_u32(tile + swz64(r, ch)));
}

// ================================================================================================
// forward
// ================================================================================================
constexpr int kFwdSmem = (kTile + 4 * kTile) * 2 + 2 * kBN * 4 + kTabSmemWords * 4;

__global__ void __launch_bounds__(128, 3) deform_attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  bf16* Qs = reinterpret_cast<bf16*>(smem);
  bf16* Ks = Qs + kTile;
  bf16* Vs = Ks + 2 * kTile;
  float* gs = reinterpret_cast<float*>(Vs + 2 * kTile);
  uint32_t* tab = reinterpret_cast<uint32_t*>(gs + 2 * kBN);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const bf16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const bf16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const bf16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
  const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
  const int ntiles = cdiv(p.n_kv, kBN);

  load_tile64(Qs, qb, p.ldq, i0, p.n, tid);
  load_tile64(Ks, kb, p.ldk, 0, p.n_kv, tid);
  load_tile64(Vs, vb, p.ldv, 0, p.n_kv, tid);
  if (tid < kBN) cp_async4(smem_u32(gs + tid), gb + min(tid, p.n_kv - 1), tid < p.n_kv);
  cp_async_commit();
  cpb_stage(tab, p.table, tid, 128);
  __syncthreads();
  const CpbView tb = cpb_view(tab);

  const int r_lo = warp * 16 + (lane >> 2);
  const float s_lo = seq_pos(i0 + r_lo, p.n), s_hi = seq_pos(i0 + r_lo + 8, p.n);
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const float sc2 = p.scale * kLog2e;

  for (int jt = 0; jt < ntiles; ++jt) {
    const int buf = jt & 1;
    if (jt + 1 < ntiles) {
      const int j1 = (jt + 1) * kBN;
      load_tile64(Ks + (buf ^ 1) * kTile, kb, p.ldk, j1, p.n_kv, tid);
      load_tile64(Vs + (buf ^ 1) * kTile, vb, p.ldv, j1, p.n_kv, tid);
      if (tid < kBN) cp_async4(smem_u32(gs + (buf ^ 1) * kBN + tid), gb + min(j1 + tid, p.n_kv - 1), j1 + tid < p.n_kv);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* Kt = Ks + buf * kTile;
    const bf16* Vt = Vs + buf * kTile;
    const float* gt = gs + buf * kBN;

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      uint32_t a[4];
      lda(a, Qs, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_nk(bb, Kt, np * 16, kc, lane);
        mma_bf16_16816(s[2 * np], a, bb[0], bb[1]);
        mma_bf16_16816(s[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    // scale + bias + mask, in the log2 domain
    const int jbase = jt * kBN;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = nt * 8 + (lane & 3) * 2 + e;
        const float gj = gt[jj];
        const bool valid = (jbase + jj) < p.n_kv;
        {
          const float t = cpb_t(s_lo - gj);
          const float4 c = tb.coef[cpb_segment(tb, t)];
          const float bias = oidx ? fmaf(c.z, t, c.w) : fmaf(c.x, t, c.y);
          const float v = valid ? fmaf(s[nt][e], sc2, bias * kLog2e) : -INFINITY;
          s[nt][e] = v;
          mx_lo = fmaxf(mx_lo, v);
        }
        {
          const float t = cpb_t(s_hi - gj);
          const float4 c = tb.coef[cpb_segment(tb, t)];
          const float bias = oidx ? fmaf(c.z, t, c.w) : fmaf(c.x, t, c.y);
          const float v = valid ? fmaf(s[nt][2 + e], sc2, bias * kLog2e) : -INFINITY;
          s[nt][2 + e] = v;
          mx_hi = fmaxf(mx_hi, v);
        }
      }
    }
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
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bb[4];
        ldb_kn(bb, Vt, kk * 16, dp, lane);
        mma_bf16_16816(o[2 * dp], a, bb[0], bb[1]);
        mma_bf16_16816(o[2 * dp + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
  // finalize
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
  const int gi_lo = i0 + r_lo, gi_hi = gi_lo + 8;
  bf16* ob = p.o + (size_t)b * p.n * p.ldo + h * kD + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (gi_lo < p.n) *reinterpret_cast<uint32_t*>(ob + (size_t)gi_lo * p.ldo + nt * 8) = pack_bf16(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    if (gi_hi < p.n) *reinterpret_cast<uint32_t*>(ob + (size_t)gi_hi * p.ldo + nt * 8) = pack_bf16(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
  if ((lane & 3) == 0) {
    float* lb = p.lse + ((size_t)b * p.H + h) * p.n;
    if (gi_lo < p.n) lb[gi_lo] = m_lo + log2f(l_lo);
    if (gi_hi < p.n) lb[gi_hi] = m_hi + log2f(l_hi);
  }
}

// ================================================================================================
// backward prep: D[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
// ================================================================================================
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                            int B, int n, int H, int ld, float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int per_lane = (H * kD) / 32;  // 16 for H = 8
  const int lanes_per_head = kD / per_lane;
  for (int row = warp; row < B * n; row += nwarps) {
    const bf16* po = o + (size_t)row * ld + lane * per_lane;
    const bf16* pd = d_o + (size_t)row * ld + lane * per_lane;
    float s = 0.f;
    for (int e = 0; e < per_lane; e += 8) {
      uint4 a = *reinterpret_cast<const uint4*>(po + e), c = *reinterpret_cast<const uint4*>(pd + e);
      const bf162* a2 = reinterpret_cast<const bf162*>(&a);
      const bf162* c2 = reinterpret_cast<const bf162*>(&c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s = fmaf(__low2float(a2[k]), __low2float(c2[k]), s);
        s = fmaf(__high2float(a2[k]), __high2float(c2[k]), s);
      }
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
constexpr int kDqSmem = (2 * kTile + 4 * kTile) * 2 + 2 * kBN * 4 + kTabSmemWords * 4;

__global__ void __launch_bounds__(128, 2) deform_attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  bf16* Qs = reinterpret_cast<bf16*>(smem);
  bf16* dOs = Qs + kTile;
  bf16* Ks = dOs + kTile;
  bf16* Vs = Ks + 2 * kTile;
  float* gs = reinterpret_cast<float*>(Vs + 2 * kTile);
  uint32_t* tab = reinterpret_cast<uint32_t*>(gs + 2 * kBN);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const bf16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const bf16* dob = p.d_o + (size_t)b * p.n * p.ldo + h * kD;
  const bf16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const bf16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
  const float* gb = p.g + (size_t)(b * G + grp) * p.n_kv;
  const int ntiles = cdiv(p.n_kv, kBN);

  load_tile64(Qs, qb, p.ldq, i0, p.n, tid);
  load_tile64(dOs, dob, p.ldo, i0, p.n, tid);
  load_tile64(Ks, kb, p.ldk, 0, p.n_kv, tid);
  load_tile64(Vs, vb, p.ldv, 0, p.n_kv, tid);
  if (tid < kBN) cp_async4(smem_u32(gs + tid), gb + min(tid, p.n_kv - 1), tid < p.n_kv);
  cp_async_commit();
  cpb_stage(tab, p.table, tid, 128);
  __syncthreads();
  const CpbView tb = cpb_view(tab);

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
    if (jt + 1 < ntiles) {
      const int j1 = (jt + 1) * kBN;
      load_tile64(Ks + (buf ^ 1) * kTile, kb, p.ldk, j1, p.n_kv, tid);
      load_tile64(Vs + (buf ^ 1) * kTile, vb, p.ldv, j1, p.n_kv, tid);
      if (tid < kBN) cp_async4(smem_u32(gs + (buf ^ 1) * kBN + tid), gb + min(j1 + tid, p.n_kv - 1), j1 + tid < p.n_kv);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* Kt = Ks + buf * kTile;
    const bf16* Vt = Vs + buf * kTile;
    const float* gt = gs + buf * kBN;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      uint32_t a[4], ad[4];
      lda(a, Qs, warp * 16, kc, lane);
      lda(ad, dOs, warp * 16, kc, lane);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        ldb_nk(bb, Kt, np * 16, kc, lane);
        mma_bf16_16816(s[2 * np], a, bb[0], bb[1]);
        mma_bf16_16816(s[2 * np + 1], a, bb[2], bb[3]);
        ldb_nk(bb, Vt, np * 16, kc, lane);
        mma_bf16_16816(dp[2 * np], ad, bb[0], bb[1]);
        mma_bf16_16816(dp[2 * np + 1], ad, bb[2], bb[3]);
      }
    }
    const int jbase = jt * kBN;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = nt * 8 + (lane & 3) * 2 + e;
        const float gj = gt[jj];
        const bool valid = (jbase + jj) < p.n_kv;
        {
          const float t = cpb_t(s_lo - gj);
          const float4 c = tb.coef[cpb_segment(tb, t)];
          const float bias = oidx ? fmaf(c.z, t, c.w) : fmaf(c.x, t, c.y);
          const float pr = valid ? exp2f(fmaf(s[nt][e], sc2, bias * kLog2e) - lse_lo) : 0.f;
          s[nt][e] = pr * (dp[nt][e] - d_lo);
        }
        {
          const float t = cpb_t(s_hi - gj);
          const float4 c = tb.coef[cpb_segment(tb, t)];
          const float bias = oidx ? fmaf(c.z, t, c.w) : fmaf(c.x, t, c.y);
          const float pr = valid ? exp2f(fmaf(s[nt][2 + e], sc2, bias * kLog2e) - lse_hi) : 0.f;
          s[nt][2 + e] = pr * (dp[nt][2 + e] - d_hi);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bb[4];
        ldb_kn(bb, Kt, kk * 16, dpi, lane);
        mma_bf16_16816(dq[2 * dpi], a, bb[0], bb[1]);
        mma_bf16_16816(dq[2 * dpi + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
  float* ob = p.dq + (size_t)b * p.n * (p.H * kD) + h * kD + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (gi_lo < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_lo * (p.H * kD) + nt * 8) = make_float2(dq[nt][0], dq[nt][1]);
    if (gi_hi < p.n) *reinterpret_cast<float2*>(ob + (size_t)gi_hi * (p.H * kD) + nt * 8) = make_float2(dq[nt][2], dq[nt][3]);
  }
}

// ================================================================================================
// backward dK / dV / dg / CPB segment sums (key-stationary, transposed orientation: rows = keys j)
// ================================================================================================
constexpr int kQT = 32;            // query rows per streamed tile
constexpr int kQTile = kQT * 64;
constexpr int kStages = 3;
constexpr int kDkvSmem = (2 * kTile + kStages * 2 * kQTile) * 2 + kStages * 2 * kQT * 4 + kTabSmemWords * 4 +
                         kCpbSegMax * 2 * 4;

__global__ void __launch_bounds__(128, 2) deform_attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  bf16* Ks = reinterpret_cast<bf16*>(smem);
  bf16* Vs = Ks + kTile;
  bf16* Qs = Vs + kTile;                                   // kStages x [32][64]
  bf16* dOs = Qs + kStages * kQTile;                       // kStages x [32][64]
  float* ls = reinterpret_cast<float*>(dOs + kStages * kQTile);  // kStages x (lse[32], D[32])
  uint32_t* tab = reinterpret_cast<uint32_t*>(ls + kStages * 2 * kQT);
  float* ssum = reinterpret_cast<float*>(tab + kTabSmemWords);   // [kCpbSegMax][2]  (A, B) for this head's output

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  const int grp = h / p.nout, oidx = h % p.nout;
  const int G = p.H / p.nout;
  const bf16* qb = p.q + (size_t)b * p.n * p.ldq + h * kD;
  const bf16* dob = p.d_o + (size_t)b * p.n * p.ldo + h * kD;
  const bf16* kb = p.k + (size_t)b * p.n_kv * p.ldk + h * kD;
  const bf16* vb = p.v + (size_t)b * p.n_kv * p.ldv + h * kD;
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

  load_tile64(Ks, kb, p.ldk, j0, p.n_kv, tid);
  load_tile64(Vs, vb, p.ldv, j0, p.n_kv, tid);
  load_stage(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_stage(1, 1);
  cp_async_commit();
  cpb_stage(tab, p.table, tid, 128);
  for (int i = tid; i < kCpbSegMax * 2; i += 128) ssum[i] = 0.f;
  __syncthreads();
  const CpbView tb = cpb_view(tab);

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

  for (int it = 0; it < ntiles; ++it) {
    const int st = it % kStages;
    if (it + 2 < ntiles) load_stage((it + 2) % kStages, it + 2);
    cp_async_commit();
    cp_async_wait<2>();
    __syncthreads();
    const bf16* Qt = Qs + st * kQTile;
    const bf16* dOt = dOs + st * kQTile;
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
      uint32_t ak[4], av[4];
      lda(ak, Ks, warp * 16, kc, lane);
      lda(av, Vs, warp * 16, kc, lane);
#pragma unroll
      for (int ip = 0; ip < 2; ++ip) {
        uint32_t bb[4];
        ldb_nk(bb, Qt, ip * 16, kc, lane);
        mma_bf16_16816(s[2 * ip], ak, bb[0], bb[1]);
        mma_bf16_16816(s[2 * ip + 1], ak, bb[2], bb[3]);
        ldb_nk(bb, dOt, ip * 16, kc, lane);
        mma_bf16_16816(dp[2 * ip], av, bb[0], bb[1]);
        mma_bf16_16816(dp[2 * ip + 1], av, bb[2], bb[3]);
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
        const float si = seq_pos(gi, p.n);
        const float lse_i = lt[ii], d_i = lt[kQT + ii];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float gj = r ? g_hi : g_lo;
          const bool valid = iv && (r ? jv_hi : jv_lo);
          const float t = cpb_t(si - gj);
          const int seg = cpb_segment(tb, t);
          const float4 c = tb.coef[seg];
          const float a_s = oidx ? c.z : c.x;
          const float bias = oidx ? fmaf(c.z, t, c.w) : fmaf(c.x, t, c.y);
          const float pr = valid ? exp2f(fmaf(s[nt][2 * r + e], sc2, bias * kLog2e) - lse_i) : 0.f;
          const float ds = pr * (dp[nt][2 * r + e] - d_i);
          pT[nt][2 * r + e] = pr;
          s[nt][2 * r + e] = ds;
          // d bias / d g_j = -a_s / (|p| + 1),  |p| + 1 = exp(|t|)
          const float dgc = -ds * a_s * exp2f(-fabsf(t) * kLog2e);
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
          run_b[r] = fmaf(ds, t, run_b[r]);
        }
      }
    }
    // dV += P^T dO ;  dK += dS^T Q      (k = i, two k-steps of 16)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t ap[4], as[4];
      ap[0] = pack_bf16(pT[2 * kk][0], pT[2 * kk][1]);
      ap[1] = pack_bf16(pT[2 * kk][2], pT[2 * kk][3]);
      ap[2] = pack_bf16(pT[2 * kk + 1][0], pT[2 * kk + 1][1]);
      ap[3] = pack_bf16(pT[2 * kk + 1][2], pT[2 * kk + 1][3]);
      as[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      as[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      as[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      as[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bb[4];
        ldb_kn(bb, dOt, kk * 16, dpi, lane);
        mma_bf16_16816(dv[2 * dpi], ap, bb[0], bb[1]);
        mma_bf16_16816(dv[2 * dpi + 1], ap, bb[2], bb[3]);
        ldb_kn(bb, Qt, kk * 16, dpi, lane);
        mma_bf16_16816(dk[2 * dpi], as, bb[0], bb[1]);
        mma_bf16_16816(dk[2 * dpi + 1], as, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
  // flush the run-length accumulators, then the CTA's segment sums
#pragma unroll
  for (int r = 0; r < 2; ++r)
    if (run_seg[r] >= 0 && (run_a[r] != 0.f || run_b[r] != 0.f)) {
      atomicAdd(ssum + 2 * run_seg[r], run_a[r]);
      atomicAdd(ssum + 2 * run_seg[r] + 1, run_b[r]);
    }
  __syncthreads();
  for (int i = tid; i < tb.nseg * 2; i += 128) {
    const float v = ssum[i];
    if (v != 0.f) atomicAdd(p.segsum + (i >> 1) * 4 + oidx * 2 + (i & 1), v);
  }
  // dg: reduce over the 4 lanes that share a key row, one atomic per (head, key)
  dg_lo += __shfl_xor_sync(0xffffffffu, dg_lo, 1);
  dg_lo += __shfl_xor_sync(0xffffffffu, dg_lo, 2);
  dg_hi += __shfl_xor_sync(0xffffffffu, dg_hi, 1);
  dg_hi += __shfl_xor_sync(0xffffffffu, dg_hi, 2);
  if ((lane & 3) == 0) {
    float* dgb = p.dg + (size_t)(b * G + grp) * p.n_kv;
    if (jv_lo) atomicAdd(dgb + gj_lo, dg_lo);
    if (jv_hi) atomicAdd(dgb + gj_hi, dg_hi);
  }
  const int ldg = p.H * kD;
  float* dkb = p.dk + (size_t)b * p.n_kv * ldg + h * kD + (lane & 3) * 2;
  float* dvb = p.dv + (size_t)b * p.n_kv * ldg + h * kD + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (jv_lo) {
      *reinterpret_cast<float2*>(dkb + (size_t)gj_lo * ldg + nt * 8) = make_float2(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
      *reinterpret_cast<float2*>(dvb + (size_t)gj_lo * ldg + nt * 8) = make_float2(dv[nt][0], dv[nt][1]);
    }
    if (jv_hi) {
      *reinterpret_cast<float2*>(dkb + (size_t)gj_hi * ldg + nt * 8) = make_float2(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
      *reinterpret_cast<float2*>(dvb + (size_t)gj_hi * ldg + nt * 8) = make_float2(dv[nt][2], dv[nt][3]);
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
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.g = g; p.table = (const uint32_t*)table;
  p.o = (bf16*)out; p.lse = lse;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.nout = heads_per_group;
  p.scale = scale;
  int rc = check_common(p);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(deform_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(cdiv(n, kBM), H, B);
  deform_attn_fwd_kernel<<<grid, 128, kFwdSmem, (cudaStream_t)stream>>>(p);
  DML_RETURN_LAUNCH();
}

int dml_deform_attn_bwd(const void* q, const void* k, const void* v, const float* g, const void* table,
                        const void* out, const void* d_out, const float* lse, int B, int H, int dim_head, int n,
                        int n_kv, int ldq, int ldk, int ldv, int ldo, int heads_per_group, float scale,
                        float* dsum_ws, float* dq, float* dk, float* dv, float* dg, float* segsum, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(q && k && v && g && table && out && d_out && lse && dsum_ws && dq && dk && dv && dg && segsum);
  if (dim_head != kD) return DML_EUNSUPPORTED;
  AttnParams p{};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.g = g; p.table = (const uint32_t*)table;
  p.o = (bf16*)out; p.lse = const_cast<float*>(lse); p.d_o = (const bf16*)d_out; p.dsum = dsum_ws;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dg = dg; p.segsum = segsum;
  p.B = B; p.H = H; p.n = n; p.n_kv = n_kv; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.nout = heads_per_group;
  p.scale = scale;
  int rc = check_common(p);
  if (rc) return rc;
  if (ldo != H * kD) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(deform_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(deform_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int G = H / heads_per_group;
  cudaError_t e = cudaMemsetAsync(dg, 0, sizeof(float) * (size_t)B * G * n_kv, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(segsum, 0, sizeof(float) * 4 * kCpbSegMax, st);
  if (e != cudaSuccess) return (int)e;
  const int rows = B * n;
  attn_bwd_prep_kernel<<<min(cdiv(rows, 8), 148 * 8), 256, 0, st>>>((const bf16*)out, (const bf16*)d_out, B, n, H, ldo, dsum_ws);
  deform_attn_bwd_dkv_kernel<<<dim3(cdiv(n_kv, kBN), H, B), 128, kDkvSmem, st>>>(p);
  deform_attn_bwd_dq_kernel<<<dim3(cdiv(n, kBM), H, B), 128, kDqSmem, st>>>(p);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
