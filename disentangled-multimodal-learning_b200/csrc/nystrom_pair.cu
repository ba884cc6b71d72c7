// HBM-bound kernels of NystromAttention (models/NystromAttention.py:74-157) on bf16-pair storage (x ~= hi + lo planes, the
// operand format of dml_pgemm), forward and backward:
//   landmark mean-pooling of q and k (:102-118),
//   row softmax of the long similarity rows q_l k^T (:125,137) and its backward,
//   depthwise 33-tap value convolution + residual add (:144-145) and its backward,
//   assembly of d(qkv) from the accumulated per-path gradients and the un-pooled landmark gradients,
// plus the fused 7x7 / 5x5 / 3x3 depthwise stencil of PPEG (models/mil.py:192-206) and its backward.
// Coalesced 8-byte (pair) / 16-byte (fp32) accesses, warp-shuffle and shared-memory reductions.
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {
namespace nyp {

// ---------------------------------------------------------------------------------------------------------------------
// landmark pooling: qkv pair [B, n_pad, ld] (q at column 0, k at column W), out pair [2 (q, k)][B, H, m, d]:
// out = mult * sum_{r < l} x[b, i l + r, .]; the front padding rows are zero rows of the buffer and count in the mean (Q8).
// grid (m, B, 2), block W / 4 threads.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void landmark_pool_kernel(const bf16* __restrict__ x, long long xplane, int ld, int W, int n_pad, int l, int m,
                                     int H, int d, float mult_q, float mult_k, bf16* __restrict__ out, long long oplane) {
  const int mi = blockIdx.x, b = blockIdx.y, which = blockIdx.z;
  const int c4 = threadIdx.x * 4;
  if (c4 >= W) return;
  const bf16* p = x + ((size_t)b * n_pad + (size_t)mi * l) * ld + which * W + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < l; ++r) {
    const float4 v = load_pair4(p + (size_t)r * ld, p + xplane + (size_t)r * ld);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float mult = which ? mult_k : mult_q;
  const int h = c4 / d, c = c4 % d;
  bf16* o = out + ((((size_t)which * gridDim.y + b) * H + h) * m + mi) * d + c;
  store_pair4(o, o + oplane, acc.x * mult, acc.y * mult, acc.z * mult, acc.w * mult);
}

// ---------------------------------------------------------------------------------------------------------------------
// long-row softmax, CTA per row: x fp32 [rows, cols] -> y pair; backward (y pair, dy fp32) -> dx pair
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int k = 0; k < nw; ++k) r = is_max ? fmaxf(r, red[k]) : r + red[k];
  return r;
}

__global__ void __launch_bounds__(512)
softmax_rows_fwd_kernel(const float* __restrict__ x, int cols, bf16* __restrict__ y, long long plane) {
  __shared__ float red[32];
  const float* p = x + (size_t)blockIdx.x * cols;
  bf16* o = y + (size_t)blockIdx.x * cols;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + c);
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  mx = block_reduce(mx, red, true);
  float s = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + c);
    s += (__expf(v.x - mx) + __expf(v.y - mx)) + (__expf(v.z - mx) + __expf(v.w - mx));
  }
  s = block_reduce(s, red, false);
  const float inv = 1.0f / s;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + c);
    store_pair4(o + c, o + plane + c, __expf(v.x - mx) * inv, __expf(v.y - mx) * inv, __expf(v.z - mx) * inv, __expf(v.w - mx) * inv);
  }
}

__global__ void __launch_bounds__(512)
softmax_rows_bwd_kernel(const bf16* __restrict__ y, long long yplane, const float* __restrict__ dy, int cols, bf16* __restrict__ dx,
                        long long dplane) {
  __shared__ float red[32];
  const size_t off = (size_t)blockIdx.x * cols;
  float s = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    const float4 a = load_pair4(y + off + c, y + yplane + off + c);
    const float4 g = *reinterpret_cast<const float4*>(dy + off + c);
    s += (a.x * g.x + a.y * g.y) + (a.z * g.z + a.w * g.w);
  }
  s = block_reduce(s, red, false);
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    const float4 a = load_pair4(y + off + c, y + yplane + off + c);
    const float4 g = *reinterpret_cast<const float4*>(dy + off + c);
    store_pair4(dx + off + c, dx + dplane + off + c, a.x * (g.x - s), a.y * (g.y - s), a.z * (g.z - s), a.w * (g.w - s));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// res_conv + residual add: y[b, i, c] = a[b, i, c] + sum_t w[h(c), t] v[b, i + t - K/2, c]   (zero padded in i), y as a pair.
// a: fp32 [B, n_pad, W] (the aggregation product, heads already merged in the columns); v: pair inside the qkv buffer.
// CTA = 64 rows x 128 columns staged with the halo in shared memory; a thread owns one column and 8-row strips.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRows = 64, kCols = 128, kKMax = 33, kHeadSlots = 4;

__global__ void __launch_bounds__(256, 2)
res_conv_fwd_kernel(const float* __restrict__ a, const bf16* __restrict__ v, long long vplane, int ldv, int col0,
                    const float* __restrict__ w, int K, int n_pad, int W, int d, bf16* __restrict__ y, long long yplane) {
  extern __shared__ float sm[];
  const int half = K / 2;
  const int i0 = blockIdx.x * kRows, cb = blockIdx.y * kCols, b = blockIdx.z;
  const int trow = kRows + K - 1;
  // stage the tile + halo: six iterations' loads go out before the first one is unpacked (12 iterations at K = 33)
  constexpr int kStageIters = ((kRows + kKMax - 1) * (kCols / 4) + 255) / 256;
#pragma unroll 6
  for (int it = 0; it < kStageIters; ++it) {
    const int idx = threadIdx.x + it * 256;
    if (idx < trow * (kCols / 4)) {
      const int r = idx / (kCols / 4), c4 = (idx % (kCols / 4)) * 4;
      const int gi = i0 + r - half;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gi >= 0 && gi < n_pad) {
        const bf16* p = v + ((size_t)b * n_pad + gi) * ldv + col0 + cb + c4;
        val = load_pair4(p, p + vplane);
      }
      *reinterpret_cast<float4*>(sm + r * kCols + c4) = val;
    }
  }
  __syncthreads();
  // a thread owns two adjacent columns (same head: d is even) and 16 rows in two 8-row strips: 4-byte pair stores
  const int c = (threadIdx.x % (kCols / 2)) * 2, rg = threadIdx.x / (kCols / 2);      // rg in 0..3
  const int col = cb + c, h = col / d;
  float wk[kKMax];
#pragma unroll
  for (int t = 0; t < kKMax; ++t) wk[t] = t < K ? w[h * K + t] : 0.f;
  for (int chunk = 0; chunk < 2; ++chunk) {
    const int r0 = rg * 16 + chunk * 8;
    float2 av[8];                  // the addend of the 8 rows: requested before the window is read, consumed after the taps
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gi = min(i0 + r0 + r, n_pad - 1);
      av[r] = __ldg(reinterpret_cast<const float2*>(a + ((size_t)b * n_pad + gi) * W + col));
    }
    float win0[8 + kKMax - 1], win1[8 + kKMax - 1];
#pragma unroll
    for (int t = 0; t < 8 + kKMax - 1; ++t) {
      const float2 v2 = (t < 8 + K - 1) ? *reinterpret_cast<const float2*>(sm + (r0 + t) * kCols + c) : make_float2(0.f, 0.f);
      win0[t] = v2.x; win1[t] = v2.y;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gi = i0 + r0 + r;
      if (gi < n_pad) {
        const size_t o = ((size_t)b * n_pad + gi) * W + col;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int t = 0; t < kKMax; ++t) { acc0 = fmaf(wk[t], win0[r + t], acc0); acc1 = fmaf(wk[t], win1[r + t], acc1); }
        acc0 += av[r].x; acc1 += av[r].y;
        uint32_t hi, lo;
        split_bf16x2(acc0, acc1, hi, lo);
        *reinterpret_cast<uint32_t*>(y + o) = hi;
        *reinterpret_cast<uint32_t*>(y + yplane + o) = lo;
      }
    }
  }
}

// backward: dv[b, i, c] = sum_t w[h, t] dy[b, i - t + K/2, c]  (stored into the gradient buffer: ld, col0);
//           dw[h, t] += sum_{b, i, c in head h} dy[b, i, c] v[b, i + t - K/2, c]
__global__ void __launch_bounds__(256, 2)
res_conv_bwd_kernel(const float* __restrict__ dy, const bf16* __restrict__ v, long long vplane, int ldv, int col0,
                    const float* __restrict__ w, int K, int n_pad, int W, int d, int H, float* __restrict__ dv, int lddv,
                    int dcol0, float* __restrict__ dw) {
  extern __shared__ float sm[];
  const int half = K / 2;
  const int i0 = blockIdx.x * kRows, cb = blockIdx.y * kCols, b = blockIdx.z;
  const int trow = kRows + K - 1;
  float* tdy = sm;
  float* tv = sm + trow * kCols;
  float* wred = tv + trow * kCols;
  constexpr int kStageIters = ((kRows + kKMax - 1) * (kCols / 4) + 255) / 256;
#pragma unroll 6
  for (int it = 0; it < kStageIters; ++it) {
    const int idx = threadIdx.x + it * 256;
    if (idx < trow * (kCols / 4)) {
      const int r = idx / (kCols / 4), c4 = (idx % (kCols / 4)) * 4;
      const int gi = i0 + r - half;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), vv = g;
      if (gi >= 0 && gi < n_pad) {
        g = __ldg(reinterpret_cast<const float4*>(dy + ((size_t)b * n_pad + gi) * W + cb + c4));
        const bf16* p = v + ((size_t)b * n_pad + gi) * ldv + col0 + cb + c4;
        vv = load_pair4(p, p + vplane);
      }
      *reinterpret_cast<float4*>(tdy + r * kCols + c4) = g;
      *reinterpret_cast<float4*>(tv + r * kCols + c4) = vv;
    }
  }
  for (int idx = threadIdx.x; idx < kHeadSlots * kKMax; idx += blockDim.x) wred[idx] = 0.f;
  __syncthreads();
  const int c = threadIdx.x % kCols, rg = threadIdx.x / kCols;
  const int col = cb + c, h = col / d;
  const int hl = (cb + c) / d - cb / d;
  float wkf[kKMax], gw[kKMax];      // wkf = the taps flipped: dv[i] = sum_t wkf[t] dy[i - half + t]
#pragma unroll
  for (int t = 0; t < kKMax; ++t) { wkf[t] = t < K ? w[h * K + (K - 1 - t)] : 0.f; gw[t] = 0.f; }
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int r0 = rg * 32 + chunk * 8;
    {      // dv: one 40-row window of dy in registers
      float wdy[8 + kKMax - 1];
#pragma unroll
      for (int t = 0; t < 8 + kKMax - 1; ++t) wdy[t] = (t < 8 + K - 1) ? tdy[(r0 + t) * kCols + c] : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int gi = i0 + r0 + r;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < kKMax; ++t) acc = fmaf(wkf[t], wdy[r + t], acc);
        if (gi < n_pad) dv[((size_t)b * n_pad + gi) * lddv + dcol0 + col] = acc;
      }
    }
    asm volatile("" ::: "memory");      // keep the two windows from being live together (register budget: two CTAs per SM)
    {      // dw: the window of v against dy of the 8 rows
      float wv[8 + kKMax - 1];
#pragma unroll
      for (int t = 0; t < 8 + kKMax - 1; ++t) wv[t] = (t < 8 + K - 1) ? tv[(r0 + t) * kCols + c] : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int gi = i0 + r0 + r;
        const float g = gi < n_pad ? tdy[(r0 + r + half) * kCols + c] : 0.f;      // dy at row gi (`half` is not a compile-time index)
#pragma unroll
        for (int t = 0; t < kKMax; ++t) gw[t] = fmaf(g, wv[r + t], gw[t]);
      }
    }
  }
  const int lane = threadIdx.x & 31;
  const bool warp_one_head = (d % 32) == 0;
#pragma unroll
  for (int t = 0; t < kKMax; ++t) {
    if (t < K) {
      if (warp_one_head) {
        const float s = warp_sum(gw[t]);
        if (lane == 0) atomicAdd(wred + hl * kKMax + t, s);
      } else {
        atomicAdd(dw + h * K + t, gw[t]);
      }
    }
  }
  __syncthreads();
  if (warp_one_head) {
    for (int idx = threadIdx.x; idx < kHeadSlots * kKMax; idx += blockDim.x) {
      const int slot = idx / kKMax, t = idx % kKMax;
      const int hh = cb / d + slot;
      if (t < K && hh < H && wred[idx] != 0.f) atomicAdd(dw + hh * K + t, wred[idx]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// d(qkv) pair [B, n_pad, 3W] from the accumulated fp32 gradients acc [B, n_pad, 3W] (q columns: gradient of the SCALED
// queries) and the landmark gradients dl fp32 [2 (q_l, k_l)][B, H, m, d]:
//   dq = scale * acc_q + (scale / l) dq_l[row / l],  dk = acc_k + (1 / l) dk_l[row / l],  dv = acc_v
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dqkv_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ dl, int B, int n_pad, int W, int l, int m, int H, int d,
                     float scale, bf16* __restrict__ out, long long plane, size_t total4) {
  const int W3 = 3 * W;
  for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * blockDim.x) {
    const size_t i = i4 * 4;
    const int col = (int)(i % W3);
    const size_t rowg = i / W3;
    const int row = (int)(rowg % n_pad), b = (int)(rowg / n_pad);
    float4 g = *reinterpret_cast<const float4*>(acc + i);
    const int which = col / W;
    if (which < 2) {
      const int cw = col - which * W, h = cw / d, c = cw % d;
      const float4 t = *reinterpret_cast<const float4*>(dl + ((((size_t)which * B + b) * H + h) * m + row / l) * d + c);
      const float f0 = which == 0 ? scale : 1.0f, f1 = f0 / (float)l;
      g = make_float4(fmaf(t.x, f1, g.x * f0), fmaf(t.y, f1, g.y * f0), fmaf(t.z, f1, g.z * f0), fmaf(t.w, f1, g.w * f0));
    }
    store_pair4(out + i, out + plane + i, g.x, g.y, g.z, g.w);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// PPEG (models/mil.py:192-206): y = x + conv7(x) + conv5(x) + conv3(x) on the side x side grid of tokens 1.., token 0 (cls)
// passes through.  The three depthwise kernels are summed by the caller into one 7x7 weight wsum [C, 49] and one bias
// bsum [C] (exact up to fp32 reassociation); flip = 1 applies the transposed stencil (the input gradient).
// x, y: [B, 1 + side^2, C] token-major.  CTA = 32 channels x an 8 x 16 pixel tile; the (8 + 6) x (16 + 6) input halo is staged
// in shared memory (a pixel's 32 channels = one coalesced 128-byte row, bank = channel: conflict free).  A thread owns one
// channel and one 16-pixel output row: per input row it reads 22 values once and slides the 7 taps over them in registers
// (9.6 shared loads per output instead of 49), weights in registers.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPTH = 8, kPTW = 16, kPHW = kPTW + 6, kPHH = kPTH + 6;

__device__ __forceinline__ void ppeg_stage_tile(const float* __restrict__ g, int side, int C, int c0, int y0, int x0, float* sm) {
  // sm[(hy * kPHW + hx) * 32 + ch] = g[pixel (y0 - 3 + hy, x0 - 3 + hx), channel c0 + ch]  (zero outside the grid): 16-byte
  // cp.async copies with zero fill, all in flight at once (a register-staged loop paid one L2 latency per halo pixel)
  for (int idx = threadIdx.x; idx < kPHH * kPHW * 8; idx += blockDim.x) {
    const int i = idx >> 3, ch = (idx & 7) * 4;
    const int hy = i / kPHW, hx = i - hy * kPHW;
    const int yy = y0 - 3 + hy, xx = x0 - 3 + hx;
    const bool ok = yy >= 0 && yy < side && xx >= 0 && xx < side && c0 + ch < C;
    const float* src = ok ? g + ((size_t)yy * side + xx) * C + c0 + ch : g;
    cp_async16(smem_u32(sm + i * 32 + ch), src, ok);
  }
  cp_async_commit();
  cp_async_wait<0>();
}

__global__ void __launch_bounds__(256)
ppeg_stencil_kernel(const float* __restrict__ x, const float* __restrict__ wsum, const float* __restrict__ bsum, int side, int C,
                    int flip, float* __restrict__ y) {
  __shared__ float sm[kPHH * kPHW * 32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, c = c0 + lane;
  const int tiles_x = cdiv(side, kPTW);
  const int ty = blockIdx.y / tiles_x, tx = blockIdx.y % tiles_x;
  const int y0 = ty * kPTH, x0 = tx * kPTW;
  const int b = blockIdx.z;
  const size_t base = (size_t)b * ((size_t)side * side + 1) * C;
  if (blockIdx.y == 0 && wp == 0 && c < C) y[base + c] = x[base + c];          // cls token
  ppeg_stage_tile(x + base + C, side, C, c0, y0, x0, sm);
  float w[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) w[t] = c < C ? __ldg(wsum + (size_t)c * 49 + (flip ? 48 - t : t)) : 0.f;
  const float bias = (c < C && !flip) ? __ldg(bsum + c) : 0.f;
  __syncthreads();
  const int oy = y0 + wp;                               // this warp's output row
  float acc[kPTW];
#pragma unroll
  for (int j = 0; j < kPTW; ++j) acc[j] = sm[((wp + 3) * kPHW + j + 3) * 32 + lane] + bias;      // identity term
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    float in[kPHW];
#pragma unroll
    for (int j = 0; j < kPHW; ++j) in[j] = sm[((wp + dy) * kPHW + j) * 32 + lane];
#pragma unroll
    for (int j = 0; j < kPTW; ++j)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) acc[j] = fmaf(w[dy * 7 + dx], in[j + dx], acc[j]);
  }
  if (c < C && oy < side) {
    float* yo = y + base + C + ((size_t)oy * side + x0) * C + c;
#pragma unroll
    for (int j = 0; j < kPTW; ++j)
      if (x0 + j < side) yo[(size_t)j * C] = acc[j];
  }
}

// weight / bias gradients of the summed stencil: dw[c, tap] = sum_p dy[p, c] x[p + off(tap), c], db[c] = sum_p dy[p, c].
// Same tiling; a CTA walks several pixel tiles (grid.y chunks) with its 49 + 1 sums in registers and flushes them once.
__global__ void __launch_bounds__(256)
ppeg_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int side, int C, int tiles_per_cta, float* __restrict__ dw,
                  float* __restrict__ db) {
  __shared__ float sm[kPHH * kPHW * 32];
  __shared__ float red[50][33];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, c = c0 + lane;
  const int tiles_x = cdiv(side, kPTW), ntiles = tiles_x * cdiv(side, kPTH);
  const int b = blockIdx.z;
  const size_t base = (size_t)b * ((size_t)side * side + 1) * C + C;
  float acc[50];
#pragma unroll
  for (int t = 0; t < 50; ++t) acc[t] = 0.f;
  const int t0 = blockIdx.y * tiles_per_cta, t1 = min(ntiles, t0 + tiles_per_cta);
  for (int tile = t0; tile < t1; ++tile) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * kPTH, x0 = tx * kPTW;
    __syncthreads();
    ppeg_stage_tile(x + base, side, C, c0, y0, x0, sm);
    __syncthreads();
    const int oy = y0 + wp;
    float g[kPTW];
#pragma unroll
    for (int j = 0; j < kPTW; ++j) {
      g[j] = (c < C && oy < side && x0 + j < side) ? dy[base + ((size_t)oy * side + x0 + j) * C + c] : 0.f;
      acc[49] += g[j];
    }
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      float in[kPHW];
#pragma unroll
      for (int j = 0; j < kPHW; ++j) in[j] = sm[((wp + d) * kPHW + j) * 32 + lane];
#pragma unroll
      for (int dx = 0; dx < 7; ++dx)
#pragma unroll
        for (int j = 0; j < kPTW; ++j) acc[d * 7 + dx] = fmaf(g[j], in[j + dx], acc[d * 7 + dx]);
    }
  }
  // reduce over the CTA's 8 warps (output rows), then one atomic per (channel, tap)
  for (int t = threadIdx.x; t < 50 * 33; t += blockDim.x) (&red[0][0])[t] = 0.f;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 50; ++t) atomicAdd(&red[t][lane], acc[t]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < 50 * 32; idx += blockDim.x) {
    const int t = idx >> 5, ln = idx & 31;
    const int cc = c0 + ln;
    if (cc < C) {
      if (t < 49) atomicAdd(dw + (size_t)cc * 49 + t, red[t][ln]);
      else atomicAdd(db + cc, red[t][ln]);
    }
  }
}

}  // namespace nyp
}  // namespace dml

extern "C" {

int dml_ny_landmark_pool(const void* qkv, long long plane_stride, int ld, int B, int n_pad, int l, int H, int d, float mult_q,
                         float mult_k, void* out, long long out_plane_stride, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(qkv && out && B > 0 && n_pad > 0 && l > 0 && H > 0 && d > 0 && (n_pad % l) == 0);
  const int W = H * d;
  if ((d % 4) || (ld % 4) || (plane_stride % 4) || (out_plane_stride % 4) || W / 4 > 1024 || ld < 2 * W) return DML_EUNSUPPORTED;
  const int m = n_pad / l;
  nyp::landmark_pool_kernel<<<dim3(m, B, 2), cdiv(W / 4, 32) * 32, 0, (cudaStream_t)stream>>>(
      (const bf16*)qkv, plane_stride, ld, W, n_pad, l, m, H, d, mult_q, mult_k, (bf16*)out, out_plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_ny_softmax_rows_fwd(const float* x, long long rows, int cols, void* y, long long plane_stride, void* stream) {
  DML_CHECK_ARG(x && y && rows > 0 && cols > 0);
  if ((cols % 4) || (plane_stride % 4) || rows > 0x7fffffffLL) return DML_EUNSUPPORTED;
  dml::nyp::softmax_rows_fwd_kernel<<<(int)rows, 512, 0, (cudaStream_t)stream>>>(x, cols, (dml::bf16*)y, plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_ny_softmax_rows_bwd(const void* y, long long y_plane_stride, const float* dy, long long rows, int cols, void* dx,
                            long long dx_plane_stride, void* stream) {
  DML_CHECK_ARG(y && dy && dx && rows > 0 && cols > 0);
  if ((cols % 4) || (y_plane_stride % 4) || (dx_plane_stride % 4) || rows > 0x7fffffffLL) return DML_EUNSUPPORTED;
  dml::nyp::softmax_rows_bwd_kernel<<<(int)rows, 512, 0, (cudaStream_t)stream>>>((const dml::bf16*)y, y_plane_stride, dy, cols,
                                                                               (dml::bf16*)dx, dx_plane_stride);
  DML_RETURN_LAUNCH();
}

static int ny_conv_check(int K, int W, int d, int ldv, int col0, long long plane) {
  if (K < 1 || K > dml::nyp::kKMax || (K & 1) == 0) return DML_EUNSUPPORTED;
  if ((W % dml::nyp::kCols) != 0 || (d % 4) != 0 || (ldv % 4) != 0 || (col0 % 4) != 0 || (plane % 4) != 0) return DML_EUNSUPPORTED;
  return DML_OK;
}

int dml_ny_res_conv_fwd(const float* a, const void* v, long long v_plane_stride, int ldv, int col0, const float* w, int K, int B,
                        int n_pad, int H, int d, void* y, long long y_plane_stride, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(a && v && w && y && B > 0 && n_pad > 0);
  const int W = H * d;
  int rc = ny_conv_check(K, W, d, ldv, col0, v_plane_stride);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (size_t)(nyp::kRows + K - 1) * nyp::kCols;
  cudaError_t e = cudaFuncSetAttribute(nyp::res_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e != cudaSuccess) return (int)e;
  nyp::res_conv_fwd_kernel<<<dim3(cdiv(n_pad, nyp::kRows), W / nyp::kCols, B), 256, smem, (cudaStream_t)stream>>>(
      a, (const bf16*)v, v_plane_stride, ldv, col0, w, K, n_pad, W, d, (bf16*)y, y_plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_ny_res_conv_bwd(const float* dy, const void* v, long long v_plane_stride, int ldv, int col0, const float* w, int K, int B,
                        int n_pad, int H, int d, float* dv, int lddv, int dcol0, float* dw, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(dy && v && w && dv && dw && B > 0 && n_pad > 0 && lddv > 0);
  const int W = H * d;
  int rc = ny_conv_check(K, W, d, ldv, col0, v_plane_stride);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)H * K, st);
  if (e != cudaSuccess) return (int)e;
  const size_t smem = sizeof(float) * ((size_t)2 * (nyp::kRows + K - 1) * nyp::kCols + nyp::kHeadSlots * nyp::kKMax);
  e = cudaFuncSetAttribute(nyp::res_conv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  if (e != cudaSuccess) return (int)e;
  nyp::res_conv_bwd_kernel<<<dim3(cdiv(n_pad, nyp::kRows), W / nyp::kCols, B), 256, smem, st>>>(
      dy, (const bf16*)v, v_plane_stride, ldv, col0, w, K, n_pad, W, d, H, dv, lddv, dcol0, dw);
  DML_RETURN_LAUNCH();
}

int dml_ny_dqkv_finalize(const float* acc, const float* dl, int B, int n_pad, int l, int H, int d, float scale, void* out,
                         long long plane_stride, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(acc && dl && out && B > 0 && n_pad > 0 && l > 0 && (n_pad % l) == 0);
  if ((d % 4) || (plane_stride % 4)) return DML_EUNSUPPORTED;
  const int W = H * d;
  const size_t total4 = (size_t)B * n_pad * 3 * W / 4;
  const int blocks = (int)min((total4 + 255) / 256, (size_t)148 * 16);
  nyp::dqkv_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(acc, dl, B, n_pad, W, l, n_pad / l, H, d, scale, (bf16*)out,
                                                                     plane_stride, total4);
  DML_RETURN_LAUNCH();
}

int dml_ppeg_stencil(const float* x, const float* wsum, const float* bsum, int B, int side, int C, int flip, float* y,
                     void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && wsum && bsum && y && B > 0 && side > 0 && C > 0);
  if ((C % 4) || ((((uintptr_t)x) | ((uintptr_t)y)) & 15)) return DML_EUNSUPPORTED;
  dim3 grid(cdiv(C, 32), cdiv(side, nyp::kPTH) * cdiv(side, nyp::kPTW), B);
  if (grid.y > 65535 || B > 65535) return DML_EUNSUPPORTED;
  nyp::ppeg_stencil_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, wsum, bsum, side, C, flip, y);
  DML_RETURN_LAUNCH();
}

int dml_ppeg_wgrad(const float* x, const float* dy, int B, int side, int C, float* dw, float* db, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && dy && dw && db && B > 0 && side > 0 && C > 0);
  if ((C % 4) || (((uintptr_t)x) & 15)) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 49, st);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)C, st)) != cudaSuccess) return (int)e;
  if (B > 65535) return DML_EUNSUPPORTED;
  const int ntiles = cdiv(side, nyp::kPTH) * cdiv(side, nyp::kPTW);
  const int ctas_y = max(1, min(ntiles, 148 * 4 / max(1, cdiv(C, 32) * B)));
  const int per = cdiv(ntiles, ctas_y);
  nyp::ppeg_wgrad_kernel<<<dim3(cdiv(C, 32), cdiv(ntiles, per), B), 256, 0, st>>>(x, dy, side, C, per, dw, db);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
