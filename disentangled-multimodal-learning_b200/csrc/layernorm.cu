// Row LayerNorm forward / backward (HBM-bound): the LayerNorms of DeformCrossTransLayer (models/DeformCrossTransMIL.py:44,66,
// dim 128, applied to both token streams) and TransLayer (models/mil.py:174,186, dim 512).
// One warp per row, the row lives in registers (D / 32 values per lane, 16-byte vector loads), statistics by warp
// shuffles, two-pass variance like torch.  Backward: dx per row; the weight / bias gradients are accumulated in
// registers over a grid-stride loop, reduced across the CTA's warps in shared memory and added to global memory with
// one atomic per (CTA, column).
#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {

template <int V>   // float4 vectors per lane: D = 128 * V
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, long long rows,
                     float eps, float* __restrict__ y, bf16* __restrict__ pair, long long plane, float* __restrict__ mean,
                     float* __restrict__ rstd) {
  constexpr int D = 128 * V;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 wv[V], bv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    wv[k] = *reinterpret_cast<const float4*>(w + k * 128 + lane * 4);
    bv[k] = *reinterpret_cast<const float4*>(b + k * 128 + lane * 4);
  }
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * D;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k] = *reinterpret_cast<const float4*>(xr + k * 128 + lane * 4);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k].x -= mu; v[k].y -= mu; v[k].z -= mu; v[k].w -= mu;
      q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
    }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 o;
      o.x = fmaf(v[k].x * rs, wv[k].x, bv[k].x);
      o.y = fmaf(v[k].y * rs, wv[k].y, bv[k].y);
      o.z = fmaf(v[k].z * rs, wv[k].z, bv[k].z);
      o.w = fmaf(v[k].w * rs, wv[k].w, bv[k].w);
      if (y) *reinterpret_cast<float4*>(y + r * D + k * 128 + lane * 4) = o;
      if (pair) {      // the GEMM that consumes the normalised rows takes them as a bf16 pair: written here, not in a pass of its own
        bf16* ph = pair + r * D + k * 128 + lane * 4;
        store_pair4(ph, ph + plane, o.x, o.y, o.z, o.w);
      }
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

template <int V>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd, long long rows,
                     float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int D = 128 * V;
  __shared__ float red[8][2 * D];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 wv[V], gw[V], gb[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    wv[k] = *reinterpret_cast<const float4*>(w + k * 128 + lane * 4);
    gw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    gb[k] = gw[k];
  }
  for (long long r = warp; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    float4 xh[V], g[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float4 xv = *reinterpret_cast<const float4*>(x + r * D + k * 128 + lane * 4);
      const float4 d = *reinterpret_cast<const float4*>(dy + r * D + k * 128 + lane * 4);
      xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[k] = make_float4(d.x * wv[k].x, d.y * wv[k].y, d.z * wv[k].z, d.w * wv[k].w);
      s1 += (g[k].x + g[k].y) + (g[k].z + g[k].w);
      s2 += (g[k].x * xh[k].x + g[k].y * xh[k].y) + (g[k].z * xh[k].z + g[k].w * xh[k].w);
      gw[k].x = fmaf(d.x, xh[k].x, gw[k].x); gw[k].y = fmaf(d.y, xh[k].y, gw[k].y);
      gw[k].z = fmaf(d.z, xh[k].z, gw[k].z); gw[k].w = fmaf(d.w, xh[k].w, gw[k].w);
      gb[k].x += d.x; gb[k].y += d.y; gb[k].z += d.z; gb[k].w += d.w;
    }
    const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 o;
      o.x = rs * (g[k].x - m1 - xh[k].x * m2);
      o.y = rs * (g[k].y - m1 - xh[k].y * m2);
      o.z = rs * (g[k].z - m1 - xh[k].z * m2);
      o.w = rs * (g[k].w - m1 - xh[k].w * m2);
      *reinterpret_cast<float4*>(dx + r * D + k * 128 + lane * 4) = o;
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    *reinterpret_cast<float4*>(&red[wib][k * 128 + lane * 4]) = gw[k];
    *reinterpret_cast<float4*>(&red[wib][D + k * 128 + lane * 4]) = gb[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][i];
    atomicAdd((i < D ? dw : db - D) + i, s);
  }
}

}  // namespace dml

extern "C" {

int dml_layernorm_fwd_pair(const float* x, const float* w, const float* b, long long rows, int D, float eps, float* y,
                           void* pair, long long plane_stride, float* mean, float* rstd, void* stream) {
  DML_CHECK_ARG(x && w && b && (y || pair) && mean && rstd && rows > 0);
  if (D != 128 && D != 256 && D != 512) return DML_EUNSUPPORTED;
  if (pair && ((((uintptr_t)pair) & 7) || (plane_stride & 3))) return DML_EINVAL;
  const int blocks = (int)min((rows + 7) / 8, (long long)148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  dml::bf16* pp = (dml::bf16*)pair;
  if (D == 128) dml::layernorm_fwd_kernel<1><<<blocks, 256, 0, st>>>(x, w, b, rows, eps, y, pp, plane_stride, mean, rstd);
  else if (D == 256) dml::layernorm_fwd_kernel<2><<<blocks, 256, 0, st>>>(x, w, b, rows, eps, y, pp, plane_stride, mean, rstd);
  else dml::layernorm_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, w, b, rows, eps, y, pp, plane_stride, mean, rstd);
  DML_RETURN_LAUNCH();
}

int dml_layernorm_fwd(const float* x, const float* w, const float* b, long long rows, int D, float eps, float* y,
                      float* mean, float* rstd, void* stream) {
  DML_CHECK_ARG(y);
  return dml_layernorm_fwd_pair(x, w, b, rows, D, eps, y, nullptr, 0, mean, rstd, stream);
}

int dml_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                      long long rows, int D, float* dx, float* dw, float* db, void* stream) {
  DML_CHECK_ARG(dy && x && w && mean && rstd && dx && dw && db && rows > 0);
  if (D != 128 && D != 256 && D != 512) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * D, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(db, 0, sizeof(float) * D, st);
  if (e != cudaSuccess) return (int)e;
  const int blocks = (int)min((rows + 7) / 8, (long long)148 * 2);
  if (D == 128) dml::layernorm_bwd_kernel<1><<<blocks, 256, 0, st>>>(dy, x, w, mean, rstd, rows, dx, dw, db);
  else if (D == 256) dml::layernorm_bwd_kernel<2><<<blocks, 256, 0, st>>>(dy, x, w, mean, rstd, rows, dx, dw, db);
  else dml::layernorm_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, x, w, mean, rstd, rows, dx, dw, db);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
