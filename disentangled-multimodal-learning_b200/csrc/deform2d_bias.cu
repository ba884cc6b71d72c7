// Continuous position bias of DeformCrossAttention2D (CPB, models/DeformableAttention2D.py:121-158, called at :302-305):
//   bias[b, g, i, j] = W3 . relu(W2 relu(W1 t + b1) + b2) + b3,   t = sign(p) log(|p| + 1),  p = grid_q[i] - vs[(b g), j]  (2-vector)
// A 2 -> 32 -> 32 -> 1 MLP per (query, key, group): 1 120 MACs each, 91 % of them the 32 x 32 layer.  With a 2-D input the exact
// piecewise-linear table of the 1-D module (cpb_table.cu) does not exist, so the layer runs as a GEMM on the tensor cores:
//   one warp = one query i, tiles of 16 keys;  H1 [16 keys x 32] is produced by the threads DIRECTLY in the A-fragment layout of
//   mma.m16n8k16 (each thread evaluates the 2 x 8 (key, neuron) elements it owns), Z2 = H1 W2^T is 8 MMAs, and the ReLU / W3
//   reduction happens on the accumulator fragments (thread-local + two shuffles).  Operands are bf16 pairs (hi + lo, the
//   arithmetic of pgemm.cu): 3 MMAs per product, fp32-class result.
// Backward (same tiling, everything in registers).  The upstream gradient g_p = dS_p of a pair is a per-row scalar, and rows of dS
// sum to zero (softmax): every parameter gradient is a cancellation-dominated sum, so rounding errors must either be tiny or cancel
// like the signal does - every backward product is fp32-class (24-bit operands: three bf16 parts).  The structure keeps that cheap:
// dH1 = g_p (M W2') with M = (Z2 > 0) a 0 / 1 mask (EXACT in one bf16 part; the accumulator layout of two n-tiles IS the A layout
// of one k-step, so M is built straight into A fragments), W2' = diag(w3) W2 pre-multiplied in the shared-memory table (three
// parts) and g_p applied in fp32 afterwards: 3 MMAs; dW2 = diag(w3) M^T (g_p H1) takes both operands through movmatrix (8 x 8
// transposes), M exact, g_p H1 in three parts: 3 MMAs; only the recompute of Z2 = H1 W2^T (its masks are discontinuous in Z2)
// needs both operands in three parts: 6 MMAs.
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

// fragment tables of W2 in shared memory, [which][component][ks][nt][reg][lane] packed bf16x2 (512 words per component):
//   which = 0 (forward, Z2 = H1 W2^T):  B[kdim = m][n = k] = W2[k][m]:  word = {W2[8nt+g][16ks+2t+8reg], W2[8nt+g][16ks+2t+8reg+1]}
//   which = 1 (backward, R = M W2'):    B[kdim = k][n = m] = W2'[k][m]: word = {W2'[16ks+2t+8reg][8nt+g], W2'[16ks+2t+8reg+1][8nt+g]}
// kComp = 2: bf16 pair (hi, lo; 16 bits); kComp = 3: (hi, mid, lo; 24 bits)
// row_scale (may be NULL): the backward table holds diag(row_scale) W2, i.e. w3[k] W2[k][m] (see bias_bwd_kernel)
template <int kComp>
__device__ __forceinline__ void build_w2_frags(const float* __restrict__ W2, uint32_t* tab, int nwhich, const float* row_scale = nullptr) {
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
      v0 = W2[k0 * kHid + 8 * nt + g];
      v1 = W2[(k0 + 1) * kHid + 8 * nt + g];
      if (row_scale) {
        v0 *= row_scale[k0];
        v1 *= row_scale[k0 + 1];
      }
    }
    uint32_t* dst = tab + which * kComp * 512 + r;
    if (kComp == 2) {
      split_bf16x2(v0, v1, dst[0], dst[512]);
    } else {
      split3_bf16x2(v0, v1, dst[0], dst[512], dst[1024]);
    }
  }
}

// neurons owned by a thread (t = lane & 3): e = 0..7 -> 2t + (e & 1) + 8 (e >> 1)
__device__ __forceinline__ int neuron_of(int t, int e) { return 2 * t + (e & 1) + 8 * (e >> 1); }

struct Consts {
  float w1x[8], w1y[8], b1[8], b2[8], w3[8];
};
__device__ __forceinline__ void load_consts(Consts& c, const float* W1, const float* b1, const float* b2, const float* W3, int t) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int nb = neuron_of(t, e);
    c.w1x[e] = W1[nb * 2];
    c.w1y[e] = W1[nb * 2 + 1];
    c.b1[e] = b1[nb];
    c.b2[e] = b2[nb];
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

// H1 of the thread's elements for rows g (h0) and g + 8 (h1), and the A fragments (hi / lo) of both k-steps
__device__ __forceinline__ void layer1(const Consts& c, float tx0, float ty0, float tx1, float ty1, float (&h0)[8], float (&h1)[8],
                                       uint32_t (&ahi)[2][4], uint32_t (&alo)[2][4]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    h0[e] = fmaxf(fmaf(c.w1x[e], tx0, fmaf(c.w1y[e], ty0, c.b1[e])), 0.f);
    h1[e] = fmaxf(fmaf(c.w1x[e], tx1, fmaf(c.w1y[e], ty1, c.b1[e])), 0.f);
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    split_bf16x2(h0[4 * ks], h0[4 * ks + 1], ahi[ks][0], alo[ks][0]);
    split_bf16x2(h1[4 * ks], h1[4 * ks + 1], ahi[ks][1], alo[ks][1]);
    split_bf16x2(h0[4 * ks + 2], h0[4 * ks + 3], ahi[ks][2], alo[ks][2]);
    split_bf16x2(h1[4 * ks + 2], h1[4 * ks + 3], ahi[ks][3], alo[ks][3]);
  }
}

// acc[nt] += A (16 x 32, two k-steps: hi / lo parts) x B fragments (hi / lo parts) held in registers (forward kernel: 32 words per
// thread, loaded once per warp): pair arithmetic, 3 MMAs per product, term-major so that consecutive MMAs hit different accumulators
__device__ __forceinline__ void mma_32x32_reg(float (&acc)[4][4], const uint32_t (&ahi)[2][4], const uint32_t (&alo)[2][4],
                                              const uint32_t (&bh)[2][4][2], const uint32_t (&bl)[2][4][2]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], alo[ks], bh[ks][nt][0], bh[ks][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], ahi[ks], bl[ks][nt][0], bl[ks][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], ahi[ks], bh[ks][nt][0], bh[ks][nt][1]);
  }
}

__device__ __forceinline__ uint32_t mask_pk(uint32_t pk, bool a, bool b) { return pk & ((a ? 0x0000ffffu : 0u) | (b ? 0xffff0000u : 0u)); }

// The backward runs every product at 24 bits: the parameter gradients are cancellation-dominated sums over ReLU-masked terms
// (rows of dS sum to zero), their condition number with respect to ANY operand perturbation is ~10^2..10^3, and the masks
// (Z2 > 0) are discontinuous in Z2 - with 16-bit operands the gradients came out 2e-3 .. 8e-3 off (measured); the reference's own
// fp32 result is 1e-4 .. 9e-4 off fp64 on these tensors.  Operands in three bf16 parts, 6 MMAs per product
// (hh, hm, mh, mm, hl, lh; the dropped terms are <= 2^-24).
__device__ __forceinline__ void split_rows3(const float (&h0)[8], const float (&h1)[8], uint32_t (&a)[3][2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    split3_bf16x2(h0[4 * ks], h0[4 * ks + 1], a[0][ks][0], a[1][ks][0], a[2][ks][0]);
    split3_bf16x2(h1[4 * ks], h1[4 * ks + 1], a[0][ks][1], a[1][ks][1], a[2][ks][1]);
    split3_bf16x2(h0[4 * ks + 2], h0[4 * ks + 3], a[0][ks][2], a[1][ks][2], a[2][ks][2]);
    split3_bf16x2(h1[4 * ks + 2], h1[4 * ks + 3], a[0][ks][3], a[1][ks][3], a[2][ks][3]);
  }
}
// acc[nt] += A (three parts) x table (three parts at tab, tab + 512, tab + 1024)
__device__ __forceinline__ void mma_32x32_x6(float (&acc)[4][4], const uint32_t (&a)[3][2][4], const uint32_t* tab, int lane) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t b[3][4][2];
#pragma unroll
    for (int cp = 0; cp < 3; ++cp)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int base = cp * 512 + ((ks * 4 + nt) * 2) * 32 + lane;
        b[cp][nt][0] = tab[base];
        b[cp][nt][1] = tab[base + 32];
      }
    // smallest terms first; term-major order keeps consecutive MMAs on different accumulators
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[2][ks], b[0][nt][0], b[0][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[0][ks], b[2][nt][0], b[2][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[1][ks], b[1][nt][0], b[1][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[1][ks], b[0][nt][0], b[0][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[0][ks], b[1][nt][0], b[1][nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[0][ks], b[0][nt][0], b[0][nt][1]);
  }
}

// acc[nt] += A (ONE exact part: a 0 / 1 mask) x table (three parts)
__device__ __forceinline__ void mma_32x32_m3(float (&acc)[4][4], const uint32_t (&a)[2][4], const uint32_t* tab, int lane) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t b[3][4][2];
#pragma unroll
    for (int cp = 0; cp < 3; ++cp)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int base = cp * 512 + ((ks * 4 + nt) * 2) * 32 + lane;
        b[cp][nt][0] = tab[base];
        b[cp][nt][1] = tab[base + 32];
      }
#pragma unroll
    for (int cp = 2; cp >= 0; --cp)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[ks], b[cp][nt][0], b[cp][nt][1]);
  }
}

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
  float* vs_s = reinterpret_cast<float*>(smem_u + 1024);    // 2 m floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int bg = blockIdx.y, n = side * side;
  build_w2_frags<2>(W2, tab, 1);
  for (int i = threadIdx.x; i < 2 * m; i += 256) vs_s[i] = vs[(size_t)bg * 2 * m + i];
  Consts c;
  load_consts(c, W1, b1, b2, W3, t);
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
  uint32_t* tab = smem_u;                                   // 3072 words: forward + backward fragments, three parts each
  float* gsum = reinterpret_cast<float*>(smem_u + 3072);    // kGradFloats (padded to 1200)
  float* vs_s = gsum + 1200;                                // 2 m floats
  float* dvs_s = vs_s + 2 * m;                              // 2 m floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int bg = blockIdx.y, n = side * side;
  build_w2_frags<3>(W2, tab, 2, W3);
  for (int i = threadIdx.x; i < 2 * m; i += 256) {
    vs_s[i] = vs[(size_t)bg * 2 * m + i];
    dvs_s[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 1200; i += 256) gsum[i] = 0.f;
  Consts c;
  load_consts(c, W1, b1, b2, W3, t);
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
      float h0[8], h1[8];
      float z[4][4];
      {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          h0[e] = fmaxf(fmaf(c.w1x[e], tx0, fmaf(c.w1y[e], ty0, c.b1[e])), 0.f);
          h1[e] = fmaxf(fmaf(c.w1x[e], tx1, fmaf(c.w1y[e], ty1, c.b1[e])), 0.f);
        }
        uint32_t a3[3][2][4];
        split_rows3(h0, h1, a3);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          z[nt][0] = z[nt][2] = c.b2[2 * nt];
          z[nt][1] = z[nt][3] = c.b2[2 * nt + 1];
        }
        mma_32x32_x6(z, a3, tab, lane);
      }
      // output layer; the ReLU mask M = (Z2 > 0) as A fragments (same element -> register mapping as H1): 0 / 1 is EXACT in one
      // bf16 part, and w3 rides in the table (R = M (diag(w3) W2)) and in the final scaling of dW2 - no operand split needed
      if (t == 0) ab3 += g0 + g1;
      uint32_t mk[2][4];
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
          mk[ks][0] = mask_pk(0x3f803f80u, m0[4 * ks], m0[4 * ks + 1]);
          mk[ks][1] = mask_pk(0x3f803f80u, m1[4 * ks], m1[4 * ks + 1]);
          mk[ks][2] = mask_pk(0x3f803f80u, m0[4 * ks + 2], m0[4 * ks + 3]);
          mk[ks][3] = mask_pk(0x3f803f80u, m1[4 * ks + 2], m1[4 * ks + 3]);
        }
      }
      // R = M W2', then dH1 = g_p R in fp32 (g_p is a per-row scalar: it never enters a 16-bit operand here)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) z[nt][0] = z[nt][1] = z[nt][2] = z[nt][3] = 0.f;
      mma_32x32_m3(z, mk, tab + 1536, lane);
      float dtx0 = 0.f, dty0 = 0.f, dtx1 = 0.f, dty1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a0 = h0[e] > 0.f ? g0 * z[e >> 1][e & 1] : 0.f;
        const float a1 = h1[e] > 0.f ? g1 * z[e >> 1][2 + (e & 1)] : 0.f;
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
        if (j0 + rr < m) atomicAdd(&dvs_s[2 * j0 + lane], -dt / (fabsf(p) + 1.f));     // p = q - vs
      }
      // dW2 / w3 += M^T (g H1): A' = transposed mask blocks (exact), B' = transposed blocks of g_p H1 in three bf16 parts
      //   A'(mt) = { T(pk0[2mt]), T(pk0[2mt+1]), T(pk1[2mt]), T(pk1[2mt+1]) };  pk0[nt] = mk[nt>>1][(nt&1)*2], pk1[nt] = mk[nt>>1][(nt&1)*2+1]
      uint32_t at[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        at[mt][0] = movmatrix_t(mk[mt][0]);
        at[mt][1] = movmatrix_t(mk[mt][2]);
        at[mt][2] = movmatrix_t(mk[mt][1]);
        at[mt][3] = movmatrix_t(mk[mt][3]);
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
          const float wa = W3[16 * mt + g], wb = W3[16 * mt + 8 + g];           // the w3 factor taken out of the mask operand
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
          gsum[kGW3 + nb] += aw3[e];
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
  const size_t smem = 1024 * 4 + (size_t)2 * m * 4;
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
  const size_t smem = 3072 * 4 + 1200 * 4 + (size_t)4 * m * 4;
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
