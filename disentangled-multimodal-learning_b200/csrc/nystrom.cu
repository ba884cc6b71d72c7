// HBM-bound kernels of NystromAttention (reference: models/NystromAttention.py:74-157):
//   landmark mean-pooling (:102-118), row softmax of the three similarity matrices (:137),
//   depthwise 33-tap value convolution fused with the head merge and the residual add (:144-149).
// fp32 storage (the pinv recurrence needs fp32/tf32), coalesced float4 accesses,
// warp-shuffle reductions, shared-memory staging for the sliding-window convolution.
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {

// ------------------------------------------------------------------------------------------------
// landmark pooling.  x: [B, n_pad, ld] (the fused qkv projection, heads packed in columns col0 + h*d + c),
// out [B, H, m, d] = mult * sum_{r < l} x[b, mi*l + r, col0 + h*d + c].   n_pad == m*l; the front
// zero-padding rows are part of the buffer and are counted in the mean (quirk Q8).
// grid (m, B), block = W/4 threads (W = H*d columns, float4 per thread).
// ------------------------------------------------------------------------------------------------
__global__ void landmark_pool_fwd_kernel(const float* __restrict__ x, int ld, int col0, int n_pad, int l, int m,
                                         int H, int d, float mult, float* __restrict__ out) {
  const int mi = blockIdx.x, b = blockIdx.y;
  const int c4 = threadIdx.x * 4;  // column within [0, H*d)
  const float* p = x + ((size_t)b * n_pad + (size_t)mi * l) * ld + col0 + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = 0;
  for (; r + 4 <= l; r += 4) {  // 4 independent loads in flight
    float4 a0 = *reinterpret_cast<const float4*>(p + (size_t)(r + 0) * ld);
    float4 a1 = *reinterpret_cast<const float4*>(p + (size_t)(r + 1) * ld);
    float4 a2 = *reinterpret_cast<const float4*>(p + (size_t)(r + 2) * ld);
    float4 a3 = *reinterpret_cast<const float4*>(p + (size_t)(r + 3) * ld);
    acc.x += (a0.x + a1.x) + (a2.x + a3.x);
    acc.y += (a0.y + a1.y) + (a2.y + a3.y);
    acc.z += (a0.z + a1.z) + (a2.z + a3.z);
    acc.w += (a0.w + a1.w) + (a2.w + a3.w);
  }
  for (; r < l; ++r) {
    float4 a = *reinterpret_cast<const float4*>(p + (size_t)r * ld);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  const int h = c4 / d, c = c4 % d;
  float* o = out + (((size_t)b * H + h) * m + mi) * d + c;
  *reinterpret_cast<float4*>(o) = make_float4(acc.x * mult, acc.y * mult, acc.z * mult, acc.w * mult);
}

// backward: dx[b, row, h*d + c] = mult * dout[b, h, row / l, c]      (dx contiguous [B, n_pad, H*d])
__global__ void landmark_pool_bwd_kernel(const float* __restrict__ dout, int n_pad, int l, int m, int H, int d,
                                         float mult, float* __restrict__ dx, size_t total4) {
  const int W = H * d;
  for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * blockDim.x) {
    const size_t i = i4 * 4;
    const int col = (int)(i % W);
    const size_t rowg = i / W;
    const int row = (int)(rowg % n_pad), b = (int)(rowg / n_pad);
    const int h = col / d, c = col % d;
    const float4 g = *reinterpret_cast<const float4*>(dout + (((size_t)b * H + h) * m + row / l) * d + c);
    *reinterpret_cast<float4*>(dx + i) = make_float4(g.x * mult, g.y * mult, g.z * mult, g.w * mult);
  }
}

// ------------------------------------------------------------------------------------------------
// row softmax (fp32).  Warp-per-row for cols <= 1024 (row held in registers, one read one write),
// CTA-per-row otherwise (three passes, L2 resident).  Rows are independent segments.
// ------------------------------------------------------------------------------------------------
template <int VPL>  // float4 vectors per lane: cols <= 128 * VPL
__global__ void __launch_bounds__(256) softmax_warp_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               size_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t row = warp; row < rows; row += nwarps) {
    const float* p = x + row * cols;
    float4 v[VPL];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < cols) {
        v[k] = *reinterpret_cast<const float4*>(p + c);
        mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
      }
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < cols) {
        v[k].x = __expf(v[k].x - mx); v[k].y = __expf(v[k].y - mx);
        v[k].z = __expf(v[k].z - mx); v[k].w = __expf(v[k].w - mx);
        s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    float* o = y + row * cols;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < cols) *reinterpret_cast<float4*>(o + c) = make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) softmax_warp_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                               float* __restrict__ dx, size_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t row = warp; row < rows; row += nwarps) {
    float4 a[VPL], g[VPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < cols) {
        a[k] = *reinterpret_cast<const float4*>(y + row * cols + c);
        g[k] = *reinterpret_cast<const float4*>(dy + row * cols + c);
        s += (a[k].x * g[k].x + a[k].y * g[k].y) + (a[k].z * g[k].z + a[k].w * g[k].w);
      }
    }
    s = warp_sum(s);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = (k * 32 + lane) * 4;
      if (c < cols)
        *reinterpret_cast<float4*>(dx + row * cols + c) =
            make_float4(a[k].x * (g[k].x - s), a[k].y * (g[k].y - s), a[k].z * (g[k].z - s), a[k].w * (g[k].w - s));
    }
  }
}

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

__global__ void __launch_bounds__(512) softmax_block_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                int cols) {
  __shared__ float red[32];
  const float* p = x + (size_t)blockIdx.x * cols;
  float* o = y + (size_t)blockIdx.x * cols;
  const bool vec = (cols & 3) == 0;
  float mx = -INFINITY;
  if (vec) {
    for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
      float4 v = *reinterpret_cast<const float4*>(p + c);
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, p[c]);
  }
  mx = block_reduce(mx, red, true);
  float s = 0.f;
  if (vec) {
    for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
      float4 v = *reinterpret_cast<const float4*>(p + c);
      v.x = __expf(v.x - mx); v.y = __expf(v.y - mx); v.z = __expf(v.z - mx); v.w = __expf(v.w - mx);
      s += (v.x + v.y) + (v.z + v.w);
      *reinterpret_cast<float4*>(o + c) = v;
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) { float e = __expf(p[c] - mx); s += e; o[c] = e; }
  }
  s = block_reduce(s, red, false);
  const float inv = 1.0f / s;
  if (vec) {
    for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
      float4 v = *reinterpret_cast<float4*>(o + c);
      *reinterpret_cast<float4*>(o + c) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) o[c] *= inv;
  }
}

__global__ void __launch_bounds__(512) softmax_block_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                float* __restrict__ dx, int cols) {
  __shared__ float red[32];
  const size_t off = (size_t)blockIdx.x * cols;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s = fmaf(y[off + c], dy[off + c], s);
  s = block_reduce(s, red, false);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dx[off + c] = y[off + c] * (dy[off + c] - s);
}

// ------------------------------------------------------------------------------------------------
// res_conv + head merge + residual add:
//   y[b, i, h*d + c] = a[b, h, i, c] + sum_t w[h, t] v[b, i + t - K/2, col0 + h*d + c]   (zero padded in i)
// v lives inside the fused qkv buffer ([B, n_pad, ld], value block at col0).  CTA = 64 rows x 128 columns,
// staged (64 + K - 1) x 128 in shared memory; each thread owns one column and 8-row strips (sliding window
// in registers -> (8 + K - 1)/8 shared loads per output).
// ------------------------------------------------------------------------------------------------
constexpr int kRcRows = 64, kRcCols = 128, kRcKMax = 33, kRcHeadSlots = 4;

__global__ void __launch_bounds__(256)
res_conv_merge_fwd_kernel(const float* __restrict__ a, const float* __restrict__ v, int ldv, int col0,
                          const float* __restrict__ w, int K, int n_pad, int H, int d, float* __restrict__ y) {
  extern __shared__ float sm[];  // [(64 + K - 1)][128] staged value rows (with halo)
  const int half = K / 2;
  const int i0 = blockIdx.x * kRcRows, cb = blockIdx.y * kRcCols, b = blockIdx.z;
  const int W = H * d;
  float* tile = sm;
  const int trow = kRcRows + K - 1;
  for (int idx = threadIdx.x; idx < trow * (kRcCols / 4); idx += blockDim.x) {
    const int r = idx / (kRcCols / 4), c4 = (idx % (kRcCols / 4)) * 4;
    const int gi = i0 + r - half;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gi >= 0 && gi < n_pad) val = *reinterpret_cast<const float4*>(v + ((size_t)b * n_pad + gi) * ldv + col0 + cb + c4);
    *reinterpret_cast<float4*>(tile + r * kRcCols + c4) = val;
  }
  __syncthreads();
  const int c = threadIdx.x % kRcCols, rg = threadIdx.x / kRcCols;  // rg in {0,1}: rows rg*32 .. rg*32+31
  const int col = cb + c, h = col / d, cc = col % d;
  float wk[kRcKMax];
#pragma unroll
  for (int t = 0; t < kRcKMax; ++t) wk[t] = t < K ? w[h * K + t] : 0.f;
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int r0 = rg * 32 + chunk * 8;
    float win[8 + kRcKMax - 1];
#pragma unroll
    for (int t = 0; t < 8 + kRcKMax - 1; ++t) win[t] = (t < 8 + K - 1) ? tile[(r0 + t) * kRcCols + c] : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gi = i0 + r0 + r;
      if (gi < n_pad) {
        float acc = a[(((size_t)b * H + h) * n_pad + gi) * d + cc];
#pragma unroll
        for (int t = 0; t < kRcKMax; ++t) acc = fmaf(wk[t], win[r + t], acc);
        y[((size_t)b * n_pad + gi) * W + col] = acc;
      }
    }
  }
}

// backward: da[b,h,i,c] = dy[b,i,h*d+c];  dv[b,i,h*d+c] = sum_t w[h,t] dy[b, i - t + K/2, h*d+c]
//           dw[h,t] += sum_{b,i,c} dy[b,i,h*d+c] v[b, i + t - K/2, ...]
// Same tiling; dy is staged with the halo, dv uses the flipped taps, dw is reduced per CTA then atomics.
__global__ void __launch_bounds__(256)
res_conv_merge_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ v, int ldv, int col0,
                          const float* __restrict__ w, int K, int n_pad, int H, int d, float* __restrict__ da,
                          float* __restrict__ dv, float* __restrict__ dw) {
  extern __shared__ float sm[];
  const int half = K / 2;
  const int i0 = blockIdx.x * kRcRows, cb = blockIdx.y * kRcCols, b = blockIdx.z;
  const int W = H * d;
  const int trow = kRcRows + K - 1;
  float* tdy = sm;                        // dy rows i0-half .. i0+64+half
  float* tv = sm + trow * kRcCols;        // v  rows i0-half .. i0+64+half
  float* wred = tv + trow * kRcCols;      // [kRcHeadSlots][K] partial dw of the heads this column block touches
  for (int idx = threadIdx.x; idx < trow * (kRcCols / 4); idx += blockDim.x) {
    const int r = idx / (kRcCols / 4), c4 = (idx % (kRcCols / 4)) * 4;
    const int gi = i0 + r - half;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), vv = g;
    if (gi >= 0 && gi < n_pad) {
      g = *reinterpret_cast<const float4*>(dy + ((size_t)b * n_pad + gi) * W + cb + c4);
      vv = *reinterpret_cast<const float4*>(v + ((size_t)b * n_pad + gi) * ldv + col0 + cb + c4);
    }
    *reinterpret_cast<float4*>(tdy + r * kRcCols + c4) = g;
    *reinterpret_cast<float4*>(tv + r * kRcCols + c4) = vv;
  }
  for (int idx = threadIdx.x; idx < kRcHeadSlots * kRcKMax; idx += blockDim.x) wred[idx] = 0.f;
  __syncthreads();
  const int c = threadIdx.x % kRcCols, rg = threadIdx.x / kRcCols;
  const int col = cb + c, h = col / d, cc = col % d;
  const int hl = (cb + c) / d - cb / d;  // head slot inside this 128-column block (heads may straddle block boundaries)
  float wk[kRcKMax], gw[kRcKMax];
#pragma unroll
  for (int t = 0; t < kRcKMax; ++t) { wk[t] = t < K ? w[h * K + t] : 0.f; gw[t] = 0.f; }
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int r0 = rg * 32 + chunk * 8;
    float wdy[8 + kRcKMax - 1], wv[8 + kRcKMax - 1];
#pragma unroll
    for (int t = 0; t < 8 + kRcKMax - 1; ++t) {
      wdy[t] = (t < 8 + K - 1) ? tdy[(r0 + t) * kRcCols + c] : 0.f;
      wv[t] = (t < 8 + K - 1) ? tv[(r0 + t) * kRcCols + c] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gi = i0 + r0 + r;
      if (gi < n_pad) {
        const float g = wdy[r + half];  // dy at row gi   (requires K odd; host checks)
        da[(((size_t)b * H + h) * n_pad + gi) * d + cc] = g;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < kRcKMax; ++t) {
          // dv[gi] = sum_t w[t] dy[gi - t + half]  -> window index (r + half) - t + half = r + K - 1 - t
          if (t < K) acc = fmaf(wk[t], wdy[r + K - 1 - t], acc);
          gw[t] = fmaf(g, wv[r + t], gw[t]);
        }
        dv[((size_t)b * n_pad + gi) * W + col] = acc;
      }
    }
  }
  // reduce dw over the CTA: warp shuffle across columns of the same head, then shared atomics
  const int lane = threadIdx.x & 31;
  const bool warp_one_head = (d % 32) == 0;
#pragma unroll
  for (int t = 0; t < kRcKMax; ++t) {
    if (t < K) {
      if (warp_one_head) {
        float s = warp_sum(gw[t]);
        if (lane == 0) atomicAdd(wred + hl * kRcKMax + t, s);
      } else {
        atomicAdd(dw + h * K + t, gw[t]);
      }
    }
  }
  __syncthreads();
  if (warp_one_head) {
    // the 128-column block covers heads cb/d .. (cb+127)/d
    for (int idx = threadIdx.x; idx < kRcHeadSlots * kRcKMax; idx += blockDim.x) {
      const int slot = idx / kRcKMax, t = idx % kRcKMax;
      const int hh = cb / d + slot;
      if (t < K && hh < H && wred[idx] != 0.f) atomicAdd(dw + hh * K + t, wred[idx]);
    }
  }
}

}  // namespace dml

extern "C" {

int dml_landmark_pool_fwd(const float* x, int ld, int col0, int B, int n_pad, int l, int H, int d, float mult,
                          float* out, void* stream) {
  DML_CHECK_ARG(x && out && B > 0 && n_pad > 0 && l > 0 && H > 0 && d > 0);
  DML_CHECK_ARG(n_pad % l == 0 && (ld % 4) == 0 && (col0 % 4) == 0 && (d % 4) == 0);
  const int W = H * d, m = n_pad / l;
  if (W / 4 > 1024) return DML_EUNSUPPORTED;
  dml::landmark_pool_fwd_kernel<<<dim3(m, B), W / 4, 0, (cudaStream_t)stream>>>(x, ld, col0, n_pad, l, m, H, d, mult, out);
  DML_RETURN_LAUNCH();
}

int dml_landmark_pool_bwd(const float* dout, int B, int n_pad, int l, int H, int d, float mult, float* dx,
                          void* stream) {
  DML_CHECK_ARG(dout && dx && B > 0 && n_pad > 0 && l > 0 && n_pad % l == 0 && (d % 4) == 0);
  const size_t total4 = (size_t)B * n_pad * H * d / 4;
  const int blocks = (int)min((total4 + 255) / 256, (size_t)148 * 16);
  dml::landmark_pool_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dout, n_pad, l, n_pad / l, H, d, mult, dx, total4);
  DML_RETURN_LAUNCH();
}

int dml_softmax_rows_fwd(const float* x, float* y, long long rows, int cols, void* stream) {
  DML_CHECK_ARG(x && y && rows > 0 && cols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  if ((cols & 3) == 0 && cols <= 1024) {
    const int blocks = (int)min((size_t)((rows + 7) / 8), (size_t)148 * 8);
    if (cols <= 128) dml::softmax_warp_fwd_kernel<1><<<blocks, 256, 0, st>>>(x, y, (size_t)rows, cols);
    else if (cols <= 256) dml::softmax_warp_fwd_kernel<2><<<blocks, 256, 0, st>>>(x, y, (size_t)rows, cols);
    else if (cols <= 512) dml::softmax_warp_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, y, (size_t)rows, cols);
    else dml::softmax_warp_fwd_kernel<8><<<blocks, 256, 0, st>>>(x, y, (size_t)rows, cols);
  } else {
    if (rows > 0x7fffffffLL) return DML_EUNSUPPORTED;
    dml::softmax_block_fwd_kernel<<<(int)rows, 512, 0, st>>>(x, y, cols);
  }
  DML_RETURN_LAUNCH();
}

int dml_softmax_rows_bwd(const float* y, const float* dy, float* dx, long long rows, int cols, void* stream) {
  DML_CHECK_ARG(y && dy && dx && rows > 0 && cols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  if ((cols & 3) == 0 && cols <= 1024) {
    const int blocks = (int)min((size_t)((rows + 7) / 8), (size_t)148 * 8);
    if (cols <= 128) dml::softmax_warp_bwd_kernel<1><<<blocks, 256, 0, st>>>(y, dy, dx, (size_t)rows, cols);
    else if (cols <= 256) dml::softmax_warp_bwd_kernel<2><<<blocks, 256, 0, st>>>(y, dy, dx, (size_t)rows, cols);
    else if (cols <= 512) dml::softmax_warp_bwd_kernel<4><<<blocks, 256, 0, st>>>(y, dy, dx, (size_t)rows, cols);
    else dml::softmax_warp_bwd_kernel<8><<<blocks, 256, 0, st>>>(y, dy, dx, (size_t)rows, cols);
  } else {
    if (rows > 0x7fffffffLL) return DML_EUNSUPPORTED;
    dml::softmax_block_bwd_kernel<<<(int)rows, 512, 0, st>>>(y, dy, dx, cols);
  }
  DML_RETURN_LAUNCH();
}

static int res_conv_check(int K, int H, int d, int ldv, int col0) {
  if (K < 1 || K > dml::kRcKMax || (K & 1) == 0) return DML_EUNSUPPORTED;
  if (((H * d) % dml::kRcCols) != 0 || (d % 4) != 0 || (ldv % 4) != 0 || (col0 % 4) != 0) return DML_EUNSUPPORTED;
  return DML_OK;
}

int dml_res_conv_merge_fwd(const float* a, const float* v, int ldv, int col0, const float* w, int K, int B, int n_pad,
                           int H, int d, float* y, void* stream) {
  DML_CHECK_ARG(a && v && w && y && B > 0 && n_pad > 0);
  int rc = res_conv_check(K, H, d, ldv, col0);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (size_t)(dml::kRcRows + K - 1) * dml::kRcCols;
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    cudaError_t e = cudaFuncSetAttribute(dml::res_conv_merge_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(dml::cdiv(n_pad, dml::kRcRows), (H * d) / dml::kRcCols, B);
  dml::res_conv_merge_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a, v, ldv, col0, w, K, n_pad, H, d, y);
  DML_RETURN_LAUNCH();
}

int dml_res_conv_merge_bwd(const float* dy, const float* v, int ldv, int col0, const float* w, int K, int B, int n_pad,
                           int H, int d, float* da, float* dv, float* dw, void* stream) {
  DML_CHECK_ARG(dy && v && w && da && dv && dw && B > 0 && n_pad > 0);
  int rc = res_conv_check(K, H, d, ldv, col0);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)H * K, st);
  if (e != cudaSuccess) return (int)e;
  const size_t smem = sizeof(float) * ((size_t)2 * (dml::kRcRows + K - 1) * dml::kRcCols + dml::kRcHeadSlots * dml::kRcKMax);
  {   // per-device attribute: set on every call (cheap, no process-global flag)
    e = cudaFuncSetAttribute(dml::res_conv_merge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(dml::cdiv(n_pad, dml::kRcRows), (H * d) / dml::kRcCols, B);
  dml::res_conv_merge_bwd_kernel<<<grid, 256, smem, st>>>(dy, v, ldv, col0, w, K, n_pad, H, d, da, dv, dw);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
