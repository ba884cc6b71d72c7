// Continuous position bias of DeformCrossAttention2D (CPB, models/DeformableAttention2D.py:121-158, called at :302-305):
//   bias[b, g, i, j] = W3 . relu(W2 relu(W1 t + b1) + b2) + b3,   t = sign(p) log(|p| + 1),  p = grid_q[i] - vs[(b g), j]  (2-vector)
// A 2 -> 32 -> 32 -> 1 MLP per (query, key, group): 1 120 MACs each, 91 % of them the 32 x 32 layer.  With a 2-D input the exact
// piecewise-linear table of the 1-D module (cpb_table.cu) does not exist, so the layer runs as a GEMM on the tensor cores:
//   one warp = one query i, tiles of 16 keys;  H1 [16 keys x 32] is produced by the threads DIRECTLY in the A-fragment layout of
//   mma.m16n8k16 (each thread evaluates the 2 x 8 (key, neuron) elements it owns), Z2 = H1 W2^T is 8 MMAs per operand-part pair,
//   and the ReLU / W3 reduction happens on the accumulator fragments (thread-local + two shuffles).
// Operand arithmetic.  An mma.sync costs the scheduler ~8.5 issue cycles on this part (measured: the MMA pipe time ADDS to the
// issue time of the other instructions), so the number of MMAs matters as much as the instruction count.  H1 and W2 are O(1)
// quantities with a known bound, so they go in as fp16 PAIRS (hi + lo = 22 bits, 3 MMAs per product: hh, hl, lh) after an exact
// power-of-two scaling that puts their maxima at 2^12 (the lo part of anything above 3e-5 of the maximum is then a normal fp16);
// the scale is undone in fp32 on the accumulators.  Quantities of unknown magnitude (anything that carries the upstream gradient)
// use bf16 parts (fp32 exponent range).
// Backward (same tiling, everything in registers).  The upstream gradient g_p = dS_p of a pair is a per-row scalar, and rows of dS
// sum to zero (softmax): every parameter gradient is a cancellation-dominated sum whose condition number with respect to ANY
// operand perturbation is 10^2..10^3, and the ReLU masks are discontinuous - every backward product is fp32-class (>= 22 bits;
// with 16-bit operands the gradients came out 2e-3..8e-3 off, measured).  The structure keeps that cheap:
//   Z2 recompute  = H1 W2^T              fp16 pairs, 3 MMAs (masks need the accuracy, not the range);
//   dH1 = g_p (M W2'), M = (Z2 > 0)      the 0 / 1 mask is EXACT in one fp16 part (the accumulator layout of two n-tiles IS the A
//                                        layout of one k-step: M is built in place), W2' = diag(w3) W2 pre-multiplied in the table
//                                        as an fp16 pair, g_p applied in fp32 afterwards: 2 MMAs;
//   dW2 = diag(w3) M^T (g_p H1)          both operands through movmatrix (8 x 8 transposes), M exact (one bf16 part), g_p H1 in
//                                        three bf16 parts: 3 MMAs.
// tcgen05 is not used here: the M = 16-key tiles are produced in registers, K = N = 32, and a TMEM round trip per tile would
// cost more than the MMAs it feeds.
#include "common.cuh"

namespace dml {
namespace {

constexpr int kHid = 32;
constexpr int kFwdQ = 4;   // queries per warp (forward)
constexpr int kBwdQ = 8;   // queries per warp (backward)

__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}

// (x0, x1) -> three packed bf16 pairs h + m + l (24 significant bits)
__device__ __forceinline__ void split3_bf16x2(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float r0 = x0 - bf16_lo_f(h), r1 = x1 - bf16_hi_f(h);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(r1), "f"(r0));
  const float s0 = r0 - bf16_lo_f(m), s1 = r1 - bf16_hi_f(m);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(s1), "f"(s0));
}

// fragment tables of W2 in shared memory, [which][part][ks][nt][reg][lane] packed fp16x2, hi then lo (512 words per part):
//   which = 0 (forward, Z2 = H1 W2^T):  B[kdim = m][n = k] = W2[k][m]:  word = {W2[8nt+g][16ks+2t+8reg], W2[8nt+g][16ks+2t+8reg+1]}
//   which = 1 (backward, R = M W2'):    B[kdim = k][n = m] = W2'[k][m]: word = {W2'[16ks+2t+8reg][8nt+g], W2'[16ks+2t+8reg+1][8nt+g]}
// with W2' = diag(w3) W2.  scale[which]: the power of two that puts the table's maximum at 2^12 (applied here, exact).
__device__ __forceinline__ void build_w2_frags(const float* __restrict__ W2, const float* __restrict__ w3, uint32_t* tab, int nwhich,
                                               const float* scale) {
  for (int i = threadIdx.x; i < 512 * nwhich; i += blockDim.x) {
    const int which = i >> 9, r = i & 511;
    const int lane = r & 31, reg = (r >> 5) & 1, nt = (r >> 6) & 3, ks = (r >> 8) & 1;
    const int g = lane >> 2, t = lane & 3;
    float v0, v1;
    if (which == 0) {
      v0 = W2[(8 * nt + g) * kHid + 16 * ks + 2 * t + 8 * reg];
      v1 = W2[(8 * nt + g) * kHid + 16 * ks + 2 * t + 8 * reg + 1];
    } else {
      const int k0 = 16 * ks + 2 * t + 8 * reg;
      v0 = W2[k0 * kHid + 8 * nt + g] * w3[k0];
      v1 = W2[(k0 + 1) * kHid + 8 * nt + g] * w3[k0 + 1];
    }
    uint32_t* dst = tab + which * 1024 + r;
    split_f16(v0 * scale[which], v1 * scale[which], dst[0], dst[512]);
  }
}

// power of two s with max * s in (2^11, 2^12] (1 for max = 0)
__device__ __forceinline__ float pow2_scale(float mx) { return mx > 0.f ? exp2f(12.f - ceilf(log2f(mx))) : 1.f; }

// scales[0] = s_w (W2), scales[1] = s_2 (W2' = diag(w3) W2), scales[2] = s_h (bound of H1); call with all threads, then sync
__device__ __forceinline__ void compute_scales(const float* W1, const float* b1, const float* W2, const float* W3, float* scales,
                                               float* red) {
  float mw = 0.f, m2 = 0.f;
  for (int i = threadIdx.x; i < kHid * kHid; i += blockDim.x) {
    const float w = fabsf(W2[i]);
    mw = fmaxf(mw, w);
    m2 = fmaxf(m2, w * fabsf(W3[i >> 5]));
  }
  mw = warp_max(mw);
  m2 = warp_max(m2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[warp] = mw;
    red[8 + warp] = m2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, hb = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a = fmaxf(a, red[w]);
      b = fmaxf(b, red[8 + w]);
    }
    // |t| <= log(3) < 1.0987 for normalised positions in [-1, 1] displaced by at most 1 (wider positions only cost headroom: the
    // scaled maximum 2^12 is 16x below the fp16 limit)
    for (int k = 0; k < kHid; ++k) hb = fmaxf(hb, (fabsf(W1[2 * k]) + fabsf(W1[2 * k + 1])) * 1.0987f + fabsf(b1[k]));
    scales[0] = pow2_scale(a);
    scales[1] = pow2_scale(b);
    scales[2] = pow2_scale(hb);
  }
  __syncthreads();
}

// neurons owned by a thread (t = lane & 3): e = 0..7 -> 2t + (e & 1) + 8 (e >> 1)
__device__ __forceinline__ int neuron_of(int t, int e) { return 2 * t + (e & 1) + 8 * (e >> 1); }

struct Consts {
  float w1x[8], w1y[8], b1[8], b2[8], w3[8];
};
// layer 1 is scaled by s_h (H1 comes out pre-scaled), b2 by s_h s_w (the scale of the Z2 accumulators); w3 stays
__device__ __forceinline__ void load_consts(Consts& c, const float* W1, const float* b1, const float* b2, const float* W3, int t, float s_h,
                                            float s_z) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int nb = neuron_of(t, e);
    c.w1x[e] = W1[nb * 2] * s_h;
    c.w1y[e] = W1[nb * 2 + 1] * s_h;
    c.b1[e] = b1[nb] * s_h;
    c.b2[e] = b2[nb] * s_z;
    c.w3[e] = W3[nb];
  }
}

// signed log of the relative position of (row = lane >> 1, component = lane & 1) of the tile; p_out = the position itself.
// kAccurate (backward): logf instead of the lg2 approximation - the ReLU masks of layer 1 are discontinuous in t, and every mask
// that differs from the reference's is a term of the cancellation-dominated gradient sums
template <bool kAccurate>
__device__ __forceinline__ float tile_t(const float* vs_s, int j0, int m, int lane, float qx, float qy, float& p_out) {
  const int j = j0 + (lane >> 1);
  const float kvc = (j < m) ? vs_s[2 * j0 + lane] : 0.f;
  const float p = ((lane & 1) ? qy : qx) - kvc;
  p_out = p;
  const float a = fabsf(p) + 1.f;
  return copysignf(kAccurate ? logf(a) : __logf(a), p) * (p != 0.f ? 1.f : 0.f);
}

// (scaled) H1 of the thread's elements for rows g (h0) and g + 8 (h1), and the fp16 A fragments (hi / lo) of both k-steps
__device__ __forceinline__ void layer1(const Consts& c, float tx0, float ty0, float tx1, float ty1, float (&h0)[8], float (&h1)[8],
                                       uint32_t (&ahi)[2][4], uint32_t (&alo)[2][4]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    h0[e] = fmaxf(fmaf(c.w1x[e], tx0, fmaf(c.w1y[e], ty0, c.b1[e])), 0.f);
    h1[e] = fmaxf(fmaf(c.w1x[e], tx1, fmaf(c.w1y[e], ty1, c.b1[e])), 0.f);
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    split_f16(h0[4 * ks], h0[4 * ks + 1], ahi[ks][0], alo[ks][0]);
    split_f16(h1[4 * ks], h1[4 * ks + 1], ahi[ks][1], alo[ks][1]);
    split_f16(h0[4 * ks + 2], h0[4 * ks + 3], ahi[ks][2], alo[ks][2]);
    split_f16(h1[4 * ks + 2], h1[4 * ks + 3], ahi[ks][3], alo[ks][3]);
  }
}

// acc[nt] += A (16 x 32, two k-steps: fp16 hi / lo) x B fragments (fp16 hi / lo) held in registers (forward kernel: 32 words per
// thread, loaded once per warp): 3 MMAs per product (hh, hl, lh), term-major so that consecutive MMAs hit different accumulators
__device__ __forceinline__ void mma_32x32_reg(float (&acc)[4][4], const uint32_t (&ahi)[2][4], const uint32_t (&alo)[2][4],
                                              const uint32_t (&bh)[2][4][2], const uint32_t (&bl)[2][4][2]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], alo[ks], bh[ks][nt][0], bh[ks][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], ahi[ks], bl[ks][nt][0], bl[ks][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], ahi[ks], bh[ks][nt][0], bh[ks][nt][1]);
  }
}

// the same with the B fragments read from the shared-memory table (hi at tab, lo at tab + 512); kParts = 2: A has hi / lo parts
// (3 MMAs), kParts = 1: A is one exact part (a 0 / 1 mask; alo unused, 2 MMAs)
template <int kParts>
__device__ __forceinline__ void mma_32x32_tab(float (&acc)[4][4], const uint32_t (&ahi)[2][4], const uint32_t (&alo)[2][4],
                                              const uint32_t* tab, int lane) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t bh[4][2], bl[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int base = ((ks * 4 + nt) * 2) * 32 + lane;
      bh[nt][0] = tab[base];
      bh[nt][1] = tab[base + 32];
      bl[nt][0] = tab[512 + base];
      bl[nt][1] = tab[512 + base + 32];
    }
    if (kParts == 2) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], alo[ks], bh[nt][0], bh[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], ahi[ks], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16_16816(acc[nt], ahi[ks], bh[nt][0], bh[nt][1]);
  }
}

__device__ __forceinline__ uint32_t mask_pk(uint32_t pk, bool a, bool b) { return pk & ((a ? 0x0000ffffu : 0u) | (b ? 0xffff0000u : 0u)); }

__device__ __forceinline__ void query_xy(int i, int side, float& qx, float& qy) {
  const int y = i / side, x = i - y * side;
  const float den = (float)max(side - 1, 1);
  qx = 2.0f * (float)x / den - 1.0f;
  qy = 2.0f * (float)y / den - 1.0f;
}

// bias [B, 8, n, m]; grid (ceil(n / (8 kFwdQ)), B * 8); dynamic shared memory: 2 m floats (vs of the group) + 1024 words
__global__ void __launch_bounds__(256) bias_fwd_kernel(const float* __restrict__ vs, const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                                                       const float* __restrict__ b3, int side, int m, float* __restrict__ bias) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* tab = smem_u;                                   // 1024 words
  float* scales = reinterpret_cast<float*>(smem_u + 1024);  // 32 floats: s_w, s_2, s_h + reduction scratch
  float* vs_s = scales + 32;                                // 2 m floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int bg = blockIdx.y, n = side * side;
  compute_scales(W1, b1, W2, W3, scales, scales + 8);
  build_w2_frags(W2, W3, tab, 1, scales);
  for (int i = threadIdx.x; i < 2 * m; i += 256) vs_s[i] = vs[(size_t)bg * 2 * m + i];
  const float s_z = scales[2] * scales[0];                  // scale of the Z2 accumulators
  Consts c;
  load_consts(c, W1, b1, b2, W3, t, scales[2], s_z);
#pragma unroll
  for (int e = 0; e < 8; ++e) c.w3[e] *= 1.f / s_z;         // the epilogue un-scales (exact: powers of two)
  const float bias3 = b3[0];
  __syncthreads();
  uint32_t bh[2][4][2], bl[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int rg = 0; rg < 2; ++rg) {
        bh[ks][nt][rg] = tab[((ks * 4 + nt) * 2 + rg) * 32 + lane];
        bl[ks][nt][rg] = tab[512 + ((ks * 4 + nt) * 2 + rg) * 32 + lane];
      }
  for (int qi = 0; qi < kFwdQ; ++qi) {
    const int i = (blockIdx.x * 8 + warp) * kFwdQ + qi;
    if (i >= n) break;
    float qx, qy;
    query_xy(i, side, qx, qy);
    float* orow = bias + ((size_t)bg * n + i) * m;
    for (int j0 = 0; j0 < m; j0 += 16) {
      float p;
      const float tt = tile_t<false>(vs_s, j0, m, lane, qx, qy, p);
      const float tx0 = __shfl_sync(0xffffffffu, tt, 2 * g), ty0 = __shfl_sync(0xffffffffu, tt, 2 * g + 1);
      const float tx1 = __shfl_sync(0xffffffffu, tt, 2 * g + 16), ty1 = __shfl_sync(0xffffffffu, tt, 2 * g + 17);
      float h0[8], h1[8];
      uint32_t ahi[2][4], alo[2][4];
      layer1(c, tx0, ty0, tx1, ty1, h0, h1, ahi, alo);
      float z[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        z[nt][0] = z[nt][2] = c.b2[2 * nt];
        z[nt][1] = z[nt][3] = c.b2[2 * nt + 1];
      }
      mma_32x32_reg(z, ahi, alo, bh, bl);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s0 = fmaf(c.w3[e], fmaxf(z[e >> 1][e & 1], 0.f), s0);
        s1 = fmaf(c.w3[e], fmaxf(z[e >> 1][2 + (e & 1)], 0.f), s1);
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (t == 0 && j0 + g < m) orow[j0 + g] = s0 + bias3;
      if (t == 1 && j0 + 8 + g < m) orow[j0 + 8 + g] = s1 + bias3;
    }
  }
}

// layout of one gradient record (DML_DA2_BIAS_GRAD_FLOATS = 1192 floats)
constexpr int kGW1 = 0, kGb1 = 64, kGW2 = 96, kGb2 = 1120, kGW3 = 1152, kGb3 = 1184, kGradFloats = 1192;

// ds [B, 8, n, m] = gradient at the bias; parts[cta][1192]; dvs [(B 8), m, 2] += (atomics)
__global__ void __launch_bounds__(256) bias_bwd_kernel(const float* __restrict__ vs, const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                                                       const float* __restrict__ ds, int side, int m, float* __restrict__ parts,
                                                       float* __restrict__ dvs) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* tab = smem_u;                                   // 2048 words: forward + backward fragments, fp16 hi / lo each
  float* scales = reinterpret_cast<float*>(smem_u + 2048);  // 32 floats
  float* gsum = scales + 32;                                // kGradFloats (padded to 1200)
  float* vs_s = gsum + 1200;                                // 2 m floats
  float* dvs_s = vs_s + 2 * m;                              // 2 m floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int bg = blockIdx.y, n = side * side;
  compute_scales(W1, b1, W2, W3, scales, scales + 8);
  build_w2_frags(W2, W3, tab, 2, scales);
  for (int i = threadIdx.x; i < 2 * m; i += 256) {
    vs_s[i] = vs[(size_t)bg * 2 * m + i];
    dvs_s[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 1200; i += 256) gsum[i] = 0.f;
  const float s_h = scales[2], s_z = scales[2] * scales[0], inv_s2 = 1.f / scales[1];
  Consts c;
  load_consts(c, W1, b1, b2, W3, t, s_h, s_z);
  __syncthreads();

  float aW2[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int k = 0; k < 4; ++k) aW2[a][b][k] = 0.f;
  float ab2[8], aw3[8], ab1[8], aw1x[8], aw1y[8], ab3 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) ab2[e] = aw3[e] = ab1[e] = aw1x[e] = aw1y[e] = 0.f;

  for (int qi = 0; qi < kBwdQ; ++qi) {
    const int i = (blockIdx.x * 8 + warp) * kBwdQ + qi;
    if (i >= n) break;
    float qx, qy;
    query_xy(i, side, qx, qy);
    const float* grow = ds + ((size_t)bg * n + i) * m;
    for (int j0 = 0; j0 < m; j0 += 16) {
      float p;
      const float tt = tile_t<true>(vs_s, j0, m, lane, qx, qy, p);
      const float tx0 = __shfl_sync(0xffffffffu, tt, 2 * g), ty0 = __shfl_sync(0xffffffffu, tt, 2 * g + 1);
      const float tx1 = __shfl_sync(0xffffffffu, tt, 2 * g + 16), ty1 = __shfl_sync(0xffffffffu, tt, 2 * g + 17);
      const float g0 = (j0 + g < m) ? grow[j0 + g] : 0.f, g1 = (j0 + 8 + g < m) ? grow[j0 + 8 + g] : 0.f;
      float h0[8], h1[8];                      // H1 x s_h
      float z[4][4];                           // Z2 x s_h s_w
      {
        uint32_t ahi[2][4], alo[2][4];
        layer1(c, tx0, ty0, tx1, ty1, h0, h1, ahi, alo);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          z[nt][0] = z[nt][2] = c.b2[2 * nt];
          z[nt][1] = z[nt][3] = c.b2[2 * nt + 1];
        }
        mma_32x32_tab<2>(z, ahi, alo, tab, lane);
      }
      // output layer; the ReLU mask M = (Z2 > 0) as A fragments (same element -> register mapping as H1): 0 / 1 is EXACT in one
      // bf16 part, and w3 rides in the table (R = M (diag(w3) W2)) and in the final scaling of dW2 - no operand split needed
      if (t == 0) ab3 += g0 + g1;
      uint32_t mk[2][4], mkb[2][4];           // fp16 ones (R = M W2') and bf16 ones (dW2 = M^T (g H1))
      {
        bool m0[8], m1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float z0 = z[e >> 1][e & 1], z1 = z[e >> 1][2 + (e & 1)];
          m0[e] = z0 > 0.f;
          m1[e] = z1 > 0.f;
          aw3[e] += g0 * fmaxf(z0, 0.f) + g1 * fmaxf(z1, 0.f);
          ab2[e] += (m0[e] ? g0 : 0.f) + (m1[e] ? g1 : 0.f);             // x w3[e] at the end
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mkb[ks][0] = mask_pk(0x3f803f80u, m0[4 * ks], m0[4 * ks + 1]);
          mkb[ks][1] = mask_pk(0x3f803f80u, m1[4 * ks], m1[4 * ks + 1]);
          mkb[ks][2] = mask_pk(0x3f803f80u, m0[4 * ks + 2], m0[4 * ks + 3]);
          mkb[ks][3] = mask_pk(0x3f803f80u, m1[4 * ks + 2], m1[4 * ks + 3]);
#pragma unroll
          for (int r = 0; r < 4; ++r) mk[ks][r] = mkb[ks][r] & 0x3c003c00u;      // bf16 1.0 = 0x3f80 -> fp16 1.0 = 0x3c00
        }
      }
      // R = M W2', then dH1 = g_p R in fp32 (g_p is a per-row scalar: it never enters a 16-bit operand here)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) z[nt][0] = z[nt][1] = z[nt][2] = z[nt][3] = 0.f;
      mma_32x32_tab<1>(z, mk, mk, tab + 1024, lane);
      const float gs0 = g0 * inv_s2, gs1 = g1 * inv_s2;                                  // un-scale W2' here
      float dtx0 = 0.f, dty0 = 0.f, dtx1 = 0.f, dty1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a0 = h0[e] > 0.f ? gs0 * z[e >> 1][e & 1] : 0.f;
        const float a1 = h1[e] > 0.f ? gs1 * z[e >> 1][2 + (e & 1)] : 0.f;
        ab1[e] += a0 + a1;
        aw1x[e] += a0 * tx0 + a1 * tx1;
        aw1y[e] += a0 * ty0 + a1 * ty1;
        dtx0 = fmaf(c.w1x[e], a0, dtx0);
        dty0 = fmaf(c.w1y[e], a0, dty0);
        dtx1 = fmaf(c.w1x[e], a1, dtx1);
        dty1 = fmaf(c.w1y[e], a1, dty1);
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        dtx0 += __shfl_xor_sync(0xffffffffu, dtx0, o);
        dty0 += __shfl_xor_sync(0xffffffffu, dty0, o);
        dtx1 += __shfl_xor_sync(0xffffffffu, dtx1, o);
        dty1 += __shfl_xor_sync(0xffffffffu, dty1, o);
      }
      {
        // back to (row = lane >> 1, component = lane & 1): the quad of row (r & 7) holds the sums
        const int rr = lane >> 1, src = 4 * (rr & 7);
        const float x0 = __shfl_sync(0xffffffffu, dtx0, src), y0 = __shfl_sync(0xffffffffu, dty0, src);
        const float x1 = __shfl_sync(0xffffffffu, dtx1, src), y1 = __shfl_sync(0xffffffffu, dty1, src);
        const float dt = rr < 8 ? ((lane & 1) ? y0 : x0) : ((lane & 1) ? y1 : x1);
        if (j0 + rr < m) atomicAdd(&dvs_s[2 * j0 + lane], -dt * (1.f / s_h) / (fabsf(p) + 1.f));     // p = q - vs; w1 was scaled
      }
      // dW2 / w3 += M^T (g H1): A' = transposed mask blocks (exact), B' = transposed blocks of g_p H1 in three bf16 parts
      //   A'(mt) = { T(pk0[2mt]), T(pk0[2mt+1]), T(pk1[2mt]), T(pk1[2mt+1]) };  pk0[nt] = mk[nt>>1][(nt&1)*2], pk1[nt] = mk[nt>>1][(nt&1)*2+1]
      uint32_t at[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        at[mt][0] = movmatrix_t(mkb[mt][0]);
        at[mt][1] = movmatrix_t(mkb[mt][2]);
        at[mt][2] = movmatrix_t(mkb[mt][1]);
        at[mt][3] = movmatrix_t(mkb[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t q0[3], q1[3], b0[3], b1[3];
        split3_bf16x2(g0 * h0[2 * nt], g0 * h0[2 * nt + 1], q0[0], q0[1], q0[2]);      // rows 0-7 of the tile, neurons 8 nt ..
        split3_bf16x2(g1 * h1[2 * nt], g1 * h1[2 * nt + 1], q1[0], q1[1], q1[2]);      // rows 8-15
#pragma unroll
        for (int cp = 0; cp < 3; ++cp) {
          b0[cp] = movmatrix_t(q0[cp]);
          b1[cp] = movmatrix_t(q1[cp]);
        }
#pragma unroll
        for (int cp = 2; cp >= 0; --cp)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(aW2[mt][nt], at[mt], b0[cp], b1[cp]);
      }
    }
  }

  // per-neuron sums: lanes with the same t hold the same neurons for different rows -> reduce over g
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ab2[e] += __shfl_xor_sync(0xffffffffu, ab2[e], o);
      aw3[e] += __shfl_xor_sync(0xffffffffu, aw3[e], o);
      ab1[e] += __shfl_xor_sync(0xffffffffu, ab1[e], o);
      aw1x[e] += __shfl_xor_sync(0xffffffffu, aw1x[e], o);
      aw1y[e] += __shfl_xor_sync(0xffffffffu, aw1y[e], o);
    }
    ab3 += __shfl_xor_sync(0xffffffffu, ab3, o);
  }
  // CTA sum in warp order (deterministic)
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float* o = gsum + kGW2 + (16 * mt + g) * kHid + 8 * nt + 2 * t;
          const float wa = W3[16 * mt + g] * (1.f / s_h), wb = W3[16 * mt + 8 + g] * (1.f / s_h);   // w3 (taken out of the mask operand), H1's scale
          o[0] += aW2[mt][nt][0] * wa;
          o[1] += aW2[mt][nt][1] * wa;
          o[8 * kHid] += aW2[mt][nt][2] * wb;
          o[8 * kHid + 1] += aW2[mt][nt][3] * wb;
        }
      if (g == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int nb = neuron_of(t, e);
          gsum[kGW1 + nb * 2] += aw1x[e];
          gsum[kGW1 + nb * 2 + 1] += aw1y[e];
          gsum[kGb1 + nb] += ab1[e];
          gsum[kGb2 + nb] += ab2[e] * c.w3[e];
          gsum[kGW3 + nb] += aw3[e] * (1.f / s_z);
        }
        if (t == 0) gsum[kGb3] += ab3;
      }
    }
    __syncthreads();
  }
  float* out = parts + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kGradFloats;
  for (int i = threadIdx.x; i < kGradFloats; i += 256) out[i] = gsum[i];
  for (int i = threadIdx.x; i < 2 * m; i += 256) atomicAdd(&dvs[(size_t)bg * 2 * m + i], dvs_s[i]);
}

// out[i] = sum_p parts[p][i]: CTA = 32 columns x 8 slices of the partial rows, slices combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ parts, int nparts, int len, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  const int per = cdiv(nparts, 8);
  const int p0 = slice * per, p1 = min(nparts, p0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < len) {
    int p = p0;
    for (; p + 3 < p1; p += 4) {
      s0 += parts[(size_t)p * len + col];
      s1 += parts[(size_t)(p + 1) * len + col];
      s2 += parts[(size_t)(p + 2) * len + col];
      s3 += parts[(size_t)(p + 3) * len + col];
    }
    for (; p < p1; ++p) s0 += parts[(size_t)p * len + col];
  }
  red[slice][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice == 0 && col < len) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[col] = t;
  }
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {

int dml_da2_bias_fwd(const float* vs, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                     int B, int side, int m, float* bias, void* stream) {
  DML_CHECK_ARG(vs && W1 && b1 && W2 && b2 && W3 && b3 && bias && B > 0 && side > 0 && m > 0 && B * 8 <= 65535);
  const size_t smem = 1024 * 4 + 128 + (size_t)2 * m * 4;
  if (smem > 200 * 1024) return DML_EUNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(bias_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n = side * side;
  bias_fwd_kernel<<<dim3(cdiv(n, 8 * kFwdQ), B * 8), 256, smem, (cudaStream_t)stream>>>(vs, W1, b1, W2, b2, W3, b3, side, m, bias);
  DML_RETURN_LAUNCH();
}

int dml_da2_bias_bwd_parts(int B, int side) { return cdiv(side * side, 8 * kBwdQ) * B * 8; }

/* grads: float[DML_DA2_BIAS_GRAD_FLOATS] = dW1 [32][2] | db1 [32] | dW2 [32][32] | db2 [32] | dW3 [32] | db3 [1] (+ pad) */
int dml_da2_bias_bwd(const float* vs, const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* ds,
                     int B, int side, int m, float* parts, float* grads, float* dvs, void* stream) {
  DML_CHECK_ARG(vs && W1 && b1 && W2 && b2 && W3 && ds && parts && grads && dvs && B > 0 && side > 0 && m > 0 && B * 8 <= 65535);
  const size_t smem = 2048 * 4 + 128 + 1200 * 4 + (size_t)4 * m * 4;
  if (smem > 200 * 1024) return DML_EUNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(bias_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st = (cudaStream_t)stream;
  const int n = side * side;
  const dim3 grid(cdiv(n, 8 * kBwdQ), B * 8);
  bias_bwd_kernel<<<grid, 256, smem, st>>>(vs, W1, b1, W2, b2, W3, ds, side, m, parts, dvs);
  reduce_rows_kernel<<<cdiv(kGradFloats, 32), 256, 0, st>>>(parts, (int)(grid.x * grid.y), kGradFloats, grads);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
