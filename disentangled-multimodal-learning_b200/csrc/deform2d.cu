// DeformCrossAttention2D (models/DeformableAttention2D.py:162-342): projections, offset net and bilinear gather
// (the position-bias MLP lives in deform2d_bias.cu, the attention core in deform2d_attn.cu).  All tensors token-major fp32:
//   x1, x2 [B, n, 128] (n = side^2), q [B, n, 512], kvf [B, m, 128], k / v [B, m, 512] (m = hk^2), attn [B, 8, n, m].
// The module is built for the one configuration the reference constructs (Modules.py:107-126, DeformCrossTransMIL.py:45-54):
// dim 128, 8 heads = 8 offset groups, dim_head 64, grouped 1x1 projections (16 -> 64 channels per group).
#include "common.cuh"

namespace dml {
namespace {

constexpr int kDim = 128, kG = 8, kCin = 16, kC = 512;

// ---------------------------------------------------------------------------------------------------------------------
// grouped 1x1 projection (to_q :248, to_k / to_v :285 with groups = 8): y[r][g*64+c] = sum_k x[r][g*16+k] W[g*64+c][k]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gproj_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, int rows,
                                                        float* __restrict__ y) {
  __shared__ __align__(16) float Ws[kC * 20];          // row stride 20 floats: conflict-free LDS.128 across consecutive rows
  for (int i = threadIdx.x; i < kC * kCin; i += 256) Ws[(i >> 4) * 20 + (i & 15)] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * kDim);
#pragma unroll 1
    for (int g = 0; g < kG; ++g) {
      const float4 xa = xr[g * 4], xb = xr[g * 4 + 1], xc = xr[g * 4 + 2], xd = xr[g * 4 + 3];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c = g * 64 + hf * 32 + lane;
        const float4* w = reinterpret_cast<const float4*>(&Ws[c * 20]);
        const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        float a = xa.x * w0.x + xa.y * w0.y + xa.z * w0.z + xa.w * w0.w;
        a += xb.x * w1.x + xb.y * w1.y + xb.z * w1.z + xb.w * w1.w;
        a += xc.x * w2.x + xc.y * w2.y + xc.z * w2.z + xc.w * w2.w;
        a += xd.x * w3.x + xd.y * w3.y + xd.z * w3.z + xd.w * w3.w;
        y[(size_t)r * kC + c] = a;
      }
    }
  }
}

// dx[r][g*16+k] (+)= sum_c dy[r][g*64+c] W[g*64+c][k]
__global__ void __launch_bounds__(256) gproj_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ W, int rows,
                                                           int accumulate, float* __restrict__ dx) {
  __shared__ __align__(16) float Wb[64 * 132];          // [c][g*16 + k] (+4 pad): lane (g, kb) reads 4 floats at lane*4
  for (int i = threadIdx.x; i < kC * kCin; i += 256) {
    const int row = i >> 4, k = i & 15;
    Wb[(row & 63) * 132 + (row >> 6) * 16 + k] = W[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)r * kC + g * 64);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 d = dyr[c4];
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w = *reinterpret_cast<const float4*>(&Wb[(c4 * 4 + u) * 132 + lane * 4]);
        acc.x += dv[u] * w.x; acc.y += dv[u] * w.y; acc.z += dv[u] * w.z; acc.w += dv[u] * w.w;
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + (size_t)r * kDim) + lane;
    if (accumulate) { const float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *o = acc;
  }
}

// dW partial of one CTA: parts[blockIdx.x][(g*64+c)*16 + k] = sum over the CTA's rows of dy[r][g*64+c] x[r][g*16+k]
__global__ void __launch_bounds__(256) gproj_bwd_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, int rows,
                                                           float* __restrict__ parts) {
  const int c0 = threadIdx.x * 2, g = c0 >> 6;
  float a0[16], a1[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a0[k] = a1[k] = 0.f;
  const int per = cdiv(rows, gridDim.x);
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (int r = r0; r < r1; ++r) {
    const float2 d = *reinterpret_cast<const float2*>(dy + (size_t)r * kC + c0);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * kDim + g * 16);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const float4 xv = xr[k4];
      a0[k4 * 4] += d.x * xv.x; a0[k4 * 4 + 1] += d.x * xv.y; a0[k4 * 4 + 2] += d.x * xv.z; a0[k4 * 4 + 3] += d.x * xv.w;
      a1[k4 * 4] += d.y * xv.x; a1[k4 * 4 + 1] += d.y * xv.y; a1[k4 * 4 + 2] += d.y * xv.z; a1[k4 * 4 + 3] += d.y * xv.w;
    }
  }
  float* o = parts + (size_t)blockIdx.x * (kC * kCin) + c0 * 16;
#pragma unroll
  for (int k = 0; k < 16; ++k) { o[k] = a0[k]; o[16 + k] = a1[k]; }
}

// out[i] = sum_p parts[p][i] in a fixed order (deterministic second stage of every partial-sum reduction of this file)
__global__ void reduce_parts_kernel(const float* __restrict__ parts, int nparts, long long len, float* __restrict__ out,
                                    int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float s = accumulate ? out[i] : 0.f;
  for (int p = 0; p < nparts; ++p) s += parts[(size_t)p * len + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// to_offsets (:208-214, :257), vgrid (:263-266), normalize_grid (:270)
// ---------------------------------------------------------------------------------------------------------------------
struct OffW {        // shared-memory copy of the offset-net weights
  float wt[36 * 64]; // depthwise kernel, tap-major [ky*ks+kx][c]
  float b[64];
  float w2[2 * 64];
};

__device__ __forceinline__ void load_offw(OffW& s, const float* wdw, const float* bdw, const float* w2, int ks) {
  for (int i = threadIdx.x; i < 64 * ks * ks; i += blockDim.x) {
    const int c = i / (ks * ks), t = i - c * ks * ks;
    s.wt[t * 64 + c] = wdw[i];
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s.b[i] = bdw[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s.w2[i] = w2[i];
}

__device__ __forceinline__ float gelu_erf(float c) { return 0.5f * c * (1.f + erff(c * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float c) {
  return 0.5f * (1.f + erff(c * 0.70710678118654752f)) + c * 0.3989422804014327f * __expf(-0.5f * c * c);
}

// the depthwise window of key (yk, xk) of group (b, g): conv value of channels 2*lane, 2*lane+1 (bias included)
__device__ __forceinline__ float2 off_conv(const float* __restrict__ q, const OffW& s, int b, int g, int yk, int xk, int side, int ks,
                                           int stride, int pad, int lane) {
  float a0 = 0.f, a1 = 0.f;
  const size_t n = (size_t)side * side;
  for (int ky = 0; ky < ks; ++ky) {
    const int y = yk * stride - pad + ky;
    if (y < 0 || y >= side) continue;
    for (int kx = 0; kx < ks; ++kx) {
      const int x = xk * stride - pad + kx;
      if (x < 0 || x >= side) continue;
      const float2 qv = *reinterpret_cast<const float2*>(q + ((size_t)b * n + (size_t)y * side + x) * kC + g * 64 + 2 * lane);
      const float2 wv = *reinterpret_cast<const float2*>(&s.wt[(ky * ks + kx) * 64 + 2 * lane]);
      a0 += qv.x * wv.x;
      a1 += qv.y * wv.y;
    }
  }
  return make_float2(a0 + s.b[2 * lane], a1 + s.b[2 * lane + 1]);
}

__global__ void __launch_bounds__(256) offsets_fwd_kernel(const float* __restrict__ q, const float* __restrict__ wdw,
                                                          const float* __restrict__ bdw, const float* __restrict__ w2, int B,
                                                          int side, int hk, int ks, int stride, float offset_scale,
                                                          float* __restrict__ vgrid, float* __restrict__ vs) {
  __shared__ OffW s;
  load_offw(s, wdw, bdw, w2, ks);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = hk * hk, pad = (ks - stride) / 2;
  const int total = B * kG * m;
  for (int it = blockIdx.x * 8 + warp; it < total; it += gridDim.x * 8) {
    const int bg = it / m, key = it - bg * m, b = bg / kG, g = bg - b * kG;
    const int yk = key / hk, xk = key - yk * hk;
    const float2 cv = off_conv(q, s, b, g, yk, xk, side, ks, stride, pad, lane);
    const float a0 = gelu_erf(cv.x), a1 = gelu_erf(cv.y);
    float o0 = a0 * s.w2[2 * lane] + a1 * s.w2[2 * lane + 1];
    float o1 = a0 * s.w2[64 + 2 * lane] + a1 * s.w2[64 + 2 * lane + 1];
    o0 = warp_sum(o0);
    o1 = warp_sum(o1);
    if (lane == 0) {
      const float vx = (float)xk + tanhf(o0) * offset_scale, vy = (float)yk + tanhf(o1) * offset_scale;
      vgrid[((size_t)bg * 2) * m + key] = vx;
      vgrid[((size_t)bg * 2 + 1) * m + key] = vy;
      const float den = (float)max(hk - 1, 1);          // square grid: rows - 1 == cols - 1 (the reference divides x by rows - 1)
      vs[((size_t)bg * m + key) * 2] = 2.0f * vx / den - 1.0f;
      vs[((size_t)bg * m + key) * 2 + 1] = 2.0f * vy / den - 1.0f;
    }
  }
}

// Per key: gradient of the offset net.  dvs [(B G), m, 2] = d loss / d vs; dvgrid_ext (may be NULL) = gradient that reached the
// returned vgrid.  Writes dconv [(B G), m, 64] (gradient at the depthwise conv output) and the CTA's weight-gradient partial
// parts[blockIdx.x][0..2304) = dWdw [64][36] (nn layout), [2304..2368) = db, [2368..2496) = dW2 [2][64].
constexpr int kOffGradFloats = 2496;
__global__ void __launch_bounds__(256) offsets_bwd_kernel(const float* __restrict__ q, const float* __restrict__ wdw,
                                                          const float* __restrict__ bdw, const float* __restrict__ w2,
                                                          const float* __restrict__ dvs, const float* __restrict__ dvgrid_ext, int B,
                                                          int side, int hk, int ks, int stride, float offset_scale,
                                                          float* __restrict__ dconv, float* __restrict__ parts) {
  __shared__ OffW s;
  __shared__ float red[8][32];
  load_offw(s, wdw, bdw, w2, ks);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = hk * hk, pad = (ks - stride) / 2, taps = ks * ks;
  const int total = B * kG * m;
  const size_t n = (size_t)side * side;
  float gw0[36], gw1[36];
#pragma unroll
  for (int t = 0; t < 36; ++t) gw0[t] = gw1[t] = 0.f;
  float gb0 = 0.f, gb1 = 0.f, g200 = 0.f, g201 = 0.f, g210 = 0.f, g211 = 0.f;
  const float den = (float)max(hk - 1, 1);
  for (int it = blockIdx.x * 8 + warp; it < total; it += gridDim.x * 8) {
    const int bg = it / m, key = it - bg * m, b = bg / kG, g = bg - b * kG;
    const int yk = key / hk, xk = key - yk * hk;
    const float2 cv = off_conv(q, s, b, g, yk, xk, side, ks, stride, pad, lane);
    const float a0 = gelu_erf(cv.x), a1 = gelu_erf(cv.y);
    float o0 = a0 * s.w2[2 * lane] + a1 * s.w2[2 * lane + 1];
    float o1 = a0 * s.w2[64 + 2 * lane] + a1 * s.w2[64 + 2 * lane + 1];
    o0 = warp_sum(o0);
    o1 = warp_sum(o1);
    float dvx = dvs[((size_t)bg * m + key) * 2] * (2.0f / den), dvy = dvs[((size_t)bg * m + key) * 2 + 1] * (2.0f / den);
    if (dvgrid_ext) {
      dvx += dvgrid_ext[((size_t)bg * 2) * m + key];
      dvy += dvgrid_ext[((size_t)bg * 2 + 1) * m + key];
    }
    const float t0 = tanhf(o0), t1 = tanhf(o1);
    const float do0 = dvx * offset_scale * (1.f - t0 * t0), do1 = dvy * offset_scale * (1.f - t1 * t1);
    g200 += do0 * a0; g201 += do0 * a1; g210 += do1 * a0; g211 += do1 * a1;
    const float dc0 = (do0 * s.w2[2 * lane] + do1 * s.w2[64 + 2 * lane]) * gelu_erf_grad(cv.x);
    const float dc1 = (do0 * s.w2[2 * lane + 1] + do1 * s.w2[64 + 2 * lane + 1]) * gelu_erf_grad(cv.y);
    gb0 += dc0;
    gb1 += dc1;
    *reinterpret_cast<float2*>(dconv + ((size_t)bg * m + key) * 64 + 2 * lane) = make_float2(dc0, dc1);
#pragma unroll
    for (int ky = 0; ky < 6; ++ky) {
      const int y = yk * stride - pad + ky;
      if (ky >= ks || y < 0 || y >= side) continue;
#pragma unroll
      for (int kx = 0; kx < 6; ++kx) {
        const int x = xk * stride - pad + kx;
        if (kx >= ks || x < 0 || x >= side) continue;
        const float2 qv = *reinterpret_cast<const float2*>(q + ((size_t)b * n + (size_t)y * side + x) * kC + g * 64 + 2 * lane);
        gw0[ky * 6 + kx] += dc0 * qv.x;
        gw1[ky * 6 + kx] += dc1 * qv.y;
      }
    }
  }
  // CTA reduction over the 8 warps, one quantity at a time (fixed order: deterministic)
  float* out = parts + (size_t)blockIdx.x * kOffGradFloats;
  auto cta_sum2 = [&](float v0, float v1, float* dst0, float* dst1) {
    __syncthreads();
    red[warp][lane] = v0;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane];
      *dst0 = t;
    }
    __syncthreads();
    red[warp][lane] = v1;
    __syncthreads();
    if (warp == 0) {
      t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane];
      *dst1 = t;
    }
  };
#pragma unroll
  for (int ky = 0; ky < 6; ++ky)
#pragma unroll
    for (int kx = 0; kx < 6; ++kx) {
      if (ky < ks && kx < ks)
        cta_sum2(gw0[ky * 6 + kx], gw1[ky * 6 + kx], &out[(2 * lane) * taps + ky * ks + kx], &out[(2 * lane + 1) * taps + ky * ks + kx]);
    }
  cta_sum2(gb0, gb1, &out[64 * taps + 2 * lane], &out[64 * taps + 2 * lane + 1]);
  cta_sum2(g200, g201, &out[64 * taps + 64 + 2 * lane], &out[64 * taps + 64 + 2 * lane + 1]);
  cta_sum2(g210, g211, &out[64 * taps + 128 + 2 * lane], &out[64 * taps + 128 + 2 * lane + 1]);
}

// Per token: dq[b][tok][g*64+c] += sum over the (at most 2 x 2) windows that contain the token of dconv * Wdw
__global__ void __launch_bounds__(256) offsets_bwd_dq_kernel(const float* __restrict__ dconv, const float* __restrict__ wdw, int B, int side,
                                                             int hk, int ks, int stride, float* __restrict__ dq) {
  __shared__ float wt[36 * 64];
  for (int i = threadIdx.x; i < 64 * ks * ks; i += blockDim.x) {
    const int c = i / (ks * ks), t = i - c * ks * ks;
    wt[t * 64 + c] = wdw[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = hk * hk, pad = (ks - stride) / 2, n = side * side;
  const int total = B * n;
  for (int it = blockIdx.x * 8 + warp; it < total; it += gridDim.x * 8) {
    const int b = it / n, tok = it - b * n, y = tok / side, x = tok - y * side;
    // windows: 0 <= y + pad - yk*stride < ks
    const int yk1 = min(hk - 1, (y + pad) / stride), xk1 = min(hk - 1, (x + pad) / stride);
    const int yk0 = max(0, (y + pad - ks + stride) / stride), xk0 = max(0, (x + pad - ks + stride) / stride);
    for (int g = 0; g < kG; ++g) {
      float a0 = 0.f, a1 = 0.f;
      for (int yk = yk0; yk <= yk1; ++yk) {
        const int ky = y + pad - yk * stride;
        if (ky < 0 || ky >= ks) continue;
        for (int xk = xk0; xk <= xk1; ++xk) {
          const int kx = x + pad - xk * stride;
          if (kx < 0 || kx >= ks) continue;
          const float2 dc = *reinterpret_cast<const float2*>(dconv + (((size_t)b * kG + g) * m + yk * hk + xk) * 64 + 2 * lane);
          const float2 wv = *reinterpret_cast<const float2*>(&wt[(ky * ks + kx) * 64 + 2 * lane]);
          a0 += dc.x * wv.x;
          a1 += dc.y * wv.y;
        }
      }
      float2* o = reinterpret_cast<float2*>(dq + (size_t)it * kC + g * 64 + 2 * lane);
      float2 p = *o;
      p.x += a0;
      p.y += a1;
      *o = p;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// bilinear gather of the grouped x2 at vs (F.grid_sample bilinear / zeros / align_corners=False, :274-277)
// ---------------------------------------------------------------------------------------------------------------------
struct Taps {
  int x0, y0;
  float fx, fy;     // ix - x0, iy - y0
};
__device__ __forceinline__ Taps bilinear_taps(float gx, float gy, int side) {
  const float ix = ((gx + 1.f) * side - 1.f) * 0.5f, iy = ((gy + 1.f) * side - 1.f) * 0.5f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  Taps t;
  // clamp the integer part far outside the image (all four taps are then out of bounds anyway)
  t.x0 = (int)fminf(fmaxf(fx0, -4.f), (float)side + 4.f);
  t.y0 = (int)fminf(fmaxf(fy0, -4.f), (float)side + 4.f);
  t.fx = ix - fx0;
  t.fy = iy - fy0;
  return t;
}

__global__ void gather_fwd_kernel(const float* __restrict__ x2, const float* __restrict__ vs, int B, int side, int m,
                                  float* __restrict__ kvf) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * m * 32;              // (b, key, g, c4)
  if (idx >= total) return;
  const int c4 = (int)(idx & 3), g = (int)((idx >> 2) & 7);
  const long long bk = idx >> 5;
  const int b = (int)(bk / m), key = (int)(bk - (long long)b * m);
  const float2 p = *reinterpret_cast<const float2*>(vs + (((size_t)b * kG + g) * m + key) * 2);
  const Taps t = bilinear_taps(p.x, p.y, side);
  const size_t n = (size_t)side * side;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = t.x0 + dx, yy = t.y0 + dy;
      if (xx < 0 || xx >= side || yy < 0 || yy >= side) continue;
      const float w = (dx ? t.fx : 1.f - t.fx) * (dy ? t.fy : 1.f - t.fy);
      const float4 v = *reinterpret_cast<const float4*>(x2 + ((size_t)b * n + (size_t)yy * side + xx) * kDim + g * 16 + c4 * 4);
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
  *reinterpret_cast<float4*>(kvf + ((size_t)b * m + key) * kDim + g * 16 + c4 * 4) = acc;
}

// dx2 (zero-initialised by the caller) += scatter of dkvf; dvs[(B G), m, 2] += gradient through the sampling position
__global__ void gather_bwd_kernel(const float* __restrict__ dkvf, const float* __restrict__ x2, const float* __restrict__ vs, int B,
                                  int side, int m, float* __restrict__ dx2, float* __restrict__ dvs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * m * 32;
  const bool live = idx < total;
  const long long id = live ? idx : total - 1;
  const int c4 = (int)(id & 3), g = (int)((id >> 2) & 7);
  const long long bk = id >> 5;
  const int b = (int)(bk / m), key = (int)(bk - (long long)b * m);
  const float2 p = *reinterpret_cast<const float2*>(vs + (((size_t)b * kG + g) * m + key) * 2);
  const Taps t = bilinear_taps(p.x, p.y, side);
  const size_t n = (size_t)side * side;
  const float4 d = *reinterpret_cast<const float4*>(dkvf + ((size_t)b * m + key) * kDim + g * 16 + c4 * 4);
  float gix = 0.f, giy = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = t.x0 + dx, yy = t.y0 + dy;
      if (xx < 0 || xx >= side || yy < 0 || yy >= side) continue;
      const float wx = dx ? t.fx : 1.f - t.fx, wy = dy ? t.fy : 1.f - t.fy;
      const size_t off = ((size_t)b * n + (size_t)yy * side + xx) * kDim + g * 16 + c4 * 4;
      const float4 v = *reinterpret_cast<const float4*>(x2 + off);
      const float dot = d.x * v.x + d.y * v.y + d.z * v.z + d.w * v.w;
      gix += (dx ? 1.f : -1.f) * wy * dot;
      giy += (dy ? 1.f : -1.f) * wx * dot;
      if (live) {
        const float w = wx * wy;
        atomicAdd(dx2 + off, w * d.x);
        atomicAdd(dx2 + off + 1, w * d.y);
        atomicAdd(dx2 + off + 2, w * d.z);
        atomicAdd(dx2 + off + 3, w * d.w);
      }
    }
  // the four c4 threads of a (b, key, g) are adjacent lanes
  gix += __shfl_xor_sync(0xffffffffu, gix, 1);
  giy += __shfl_xor_sync(0xffffffffu, giy, 1);
  gix += __shfl_xor_sync(0xffffffffu, gix, 2);
  giy += __shfl_xor_sync(0xffffffffu, giy, 2);
  if (live && c4 == 0) {
    float* o = dvs + (((size_t)b * kG + g) * m + key) * 2;
    o[0] += gix * (0.5f * side);
    o[1] += giy * (0.5f * side);
  }
}

inline int grid_for(long long items, int per_cta, int cap) {
  long long g = (items + per_cta - 1) / per_cta;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {

int dml_da2_kv_side(int side, int ksize, int stride) {
  const int pad = (ksize - stride) / 2;
  const int v = (side + 2 * pad - ksize);
  return v < 0 ? 0 : v / stride + 1;
}

int dml_da2_gproj_fwd(const float* x, const float* W, long long rows, float* y, void* stream) {
  DML_CHECK_ARG(x && W && y && rows > 0 && rows < (1ll << 31));
  gproj_fwd_kernel<<<grid_for(rows, 8, 148 * 4), 256, 0, (cudaStream_t)stream>>>(x, W, (int)rows, y);
  DML_RETURN_LAUNCH();
}

int dml_da2_gproj_parts(long long rows) { return grid_for(rows, 64, 148); }

int dml_da2_gproj_bwd(const float* dy, const float* x, const float* W, long long rows, int accumulate_dx, float* dx, float* parts,
                      float* dW, void* stream) {
  DML_CHECK_ARG(dy && x && W && rows > 0 && rows < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) gproj_bwd_dx_kernel<<<grid_for(rows, 8, 148 * 4), 256, 0, st>>>(dy, W, (int)rows, accumulate_dx, dx);
  if (dW) {
    DML_CHECK_ARG(parts);
    const int np = dml_da2_gproj_parts(rows);
    gproj_bwd_dw_kernel<<<np, 256, 0, st>>>(dy, x, (int)rows, parts);
    reduce_parts_kernel<<<cdiv(kC * kCin, 256), 256, 0, st>>>(parts, np, kC * kCin, dW, 0);
  }
  DML_RETURN_LAUNCH();
}

int dml_da2_reduce_parts(const float* parts, int nparts, long long len, int accumulate, float* out, void* stream) {
  DML_CHECK_ARG(parts && out && nparts > 0 && len > 0);
  reduce_parts_kernel<<<(unsigned)((len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parts, nparts, len, out, accumulate);
  DML_RETURN_LAUNCH();
}

int dml_da2_offsets_fwd(const float* q, const float* wdw, const float* bdw, const float* w2, int B, int side, int ksize, int stride,
                        float offset_scale, float* vgrid, float* vs, void* stream) {
  DML_CHECK_ARG(q && wdw && bdw && w2 && vgrid && vs && B > 0 && side > 0);
  if (ksize < stride || ksize > 6 || ((ksize - stride) & 1)) return DML_EUNSUPPORTED;
  const int hk = dml_da2_kv_side(side, ksize, stride);
  if (hk < 1) return DML_EUNSUPPORTED;
  offsets_fwd_kernel<<<grid_for((long long)B * kG * hk * hk, 8, 148 * 4), 256, 0, (cudaStream_t)stream>>>(q, wdw, bdw, w2, B, side, hk, ksize,
                                                                                                          stride, offset_scale, vgrid, vs);
  DML_RETURN_LAUNCH();
}

int dml_da2_offsets_parts(int B, int side, int ksize, int stride) {
  const int hk = dml_da2_kv_side(side, ksize, stride);
  return grid_for((long long)B * kG * hk * hk, 32, 148);
}

/* grads: float[2496] = dWdw [64][ks*ks] | db [64] | dW2 [2][64] (for ks = 6; in general 64*ks*ks + 192 floats) */
int dml_da2_offsets_bwd(const float* q, const float* wdw, const float* bdw, const float* w2, const float* dvs, const float* dvgrid_ext,
                        int B, int side, int ksize, int stride, float offset_scale, float* dconv, float* parts, float* grads, float* dq,
                        void* stream) {
  DML_CHECK_ARG(q && wdw && bdw && w2 && dvs && dconv && parts && grads && dq && B > 0 && side > 0);
  if (ksize < stride || ksize > 6 || ((ksize - stride) & 1)) return DML_EUNSUPPORTED;
  const int hk = dml_da2_kv_side(side, ksize, stride);
  if (hk < 1) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int np = dml_da2_offsets_parts(B, side, ksize, stride);
  cudaMemsetAsync(parts, 0, (size_t)np * kOffGradFloats * sizeof(float), st);
  offsets_bwd_kernel<<<np, 256, 0, st>>>(q, wdw, bdw, w2, dvs, dvgrid_ext, B, side, hk, ksize, stride, offset_scale, dconv, parts);
  const int len = 64 * ksize * ksize + 192;
  // the partial rows are kOffGradFloats apart; the first `len` floats of each are live
  reduce_parts_kernel<<<cdiv(kOffGradFloats, 256), 256, 0, st>>>(parts, np, kOffGradFloats, grads, 0);
  (void)len;
  offsets_bwd_dq_kernel<<<grid_for((long long)B * side * side, 8, 148 * 8), 256, 0, st>>>(dconv, wdw, B, side, hk, ksize, stride, dq);
  DML_RETURN_LAUNCH();
}

int dml_da2_gather_fwd(const float* x2, const float* vs, int B, int side, int m, float* kvf, void* stream) {
  DML_CHECK_ARG(x2 && vs && kvf && B > 0 && side > 0 && m > 0);
  const long long total = (long long)B * m * 32;
  gather_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x2, vs, B, side, m, kvf);
  DML_RETURN_LAUNCH();
}

int dml_da2_gather_bwd(const float* dkvf, const float* x2, const float* vs, int B, int side, int m, float* dx2, float* dvs, void* stream) {
  DML_CHECK_ARG(dkvf && x2 && vs && dx2 && dvs && B > 0 && side > 0 && m > 0);
  const long long total = (long long)B * m * 32;
  gather_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dkvf, x2, vs, B, side, m, dx2, dvs);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
