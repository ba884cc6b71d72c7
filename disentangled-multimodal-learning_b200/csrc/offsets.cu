// Offset prediction and the (degenerate) key/value gather of DeformCrossAttention1D.
//
// Reference: to_offsets Sequential (DeformableAttention1D.py:139-146), vgrid/normalize_grid (:186-188, :45-48),
// grid_sample_1d (:36-43).  HBM-bound kernels: one warp per (batch*group, key j), lanes across channels so
// every global access is a coalesced 16-B (fp32) or 8-B (bf16) vector per lane; reductions by warp shuffles.
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {

constexpr int kMaxTaps = 8;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ void load4(const h16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

// ------------------------------------------------------------------------------------------------
// offsets forward: q [B, n, C] fp16 (token-major, UNSCALED queries), groups G, Cg = C/G = 128 channels/group.
//   conv[c] = b0[c] + sum_t w0[c,t] q[b, stride*j - pad + t, g*Cg + c]     (zero padding)
//   u = sum_c w2[c] gelu(conv[c]);  off = tanh(u) * offset_scale;  vgrid = j + off;
//   g = 2 vgrid / max(n_kv - 1, 1) - 1
// One warp per (b, g, j); lane owns 4 consecutive channels (Cg == 128).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
offsets_fwd_kernel(const h16* __restrict__ q, const float* __restrict__ w0, const float* __restrict__ b0,
                   const float* __restrict__ w2, int B, int n, int C, int G, int ks, int stride, int pad, int n_kv,
                   float offset_scale, float* __restrict__ vgrid, float* __restrict__ gnorm) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Cg = C / G;
  const int c0 = lane * 4;
  float w[4][kMaxTaps], bb[4], ww2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bb[e] = b0[c0 + e];
    ww2[e] = w2[c0 + e];
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) w[e][t] = t < ks ? w0[(c0 + e) * ks + t] : 0.f;
  }
  const float gden = (float)max(n_kv - 1, 1);
  const int total = B * G * n_kv;
  for (int item = warp; item < total; item += nwarps) {
    const int j = item % n_kv, bg = item / n_kv, b = bg / G, g = bg % G;
    float acc[4] = {bb[0], bb[1], bb[2], bb[3]};
    const h16* base = q + (size_t)b * n * C + g * Cg + c0;
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) {
      const int p = stride * j - pad + t;
      if (t < ks && p >= 0 && p < n) {
        float v[4];
        load4(base + (size_t)p * C, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(w[e][t], v[e], acc[e]);
      }
    }
    float u = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) u = fmaf(ww2[e], gelu_erf(acc[e]), u);
    u = warp_sum(u);
    if (lane == 0) {
      const float off = tanhf(u) * offset_scale;
      const float vg = (float)j + off;
      vgrid[item] = vg;
      gnorm[item] = (2.0f * vg) / gden - 1.0f;  // same operation order as normalize_grid (:45-48)
    }
  }
}

// offsets backward, pass A: per (b,g,j) recompute conv/gelu/u, then
//   du = d_off * scale * (1 - tanh(u)^2);  dy[c] = du * w2[c] * gelu'(conv[c])   -> dy [(B G), n_kv, Cg] (fp32)
//   dw2[c] += du * gelu(conv[c]);  db0[c] += dy[c];  dw0[c,t] += dy[c] * q[p_t, c]
// Parameter gradients are accumulated in registers over a grid-stride loop, reduced across the CTA's
// warps in shared memory and added to global with one atomic per (CTA, element).
// wgrad layout: dw0[Cg*ks] | db0[Cg] | dw2[Cg]   (zeroed by the host wrapper)
__global__ void __launch_bounds__(256)
offsets_bwd_kernel(const h16* __restrict__ q, const float* __restrict__ w0, const float* __restrict__ b0,
                   const float* __restrict__ w2, const float* __restrict__ d_off, int B, int n, int C, int G, int ks,
                   int stride, int pad, int n_kv, float offset_scale, float* __restrict__ dy,
                   float* __restrict__ wgrad) {
  __shared__ float red[8][128 * (kMaxTaps + 2)];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Cg = C / G;
  const int c0 = lane * 4;
  float w[4][kMaxTaps], bb[4], ww2[4];
  float gw0[4][kMaxTaps], gb0[4], gw2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bb[e] = b0[c0 + e];
    ww2[e] = w2[c0 + e];
    gb0[e] = 0.f;
    gw2[e] = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) {
      w[e][t] = t < ks ? w0[(c0 + e) * ks + t] : 0.f;
      gw0[e][t] = 0.f;
    }
  }
  const int total = B * G * n_kv;
  for (int item = warp; item < total; item += nwarps) {
    const int j = item % n_kv, bg = item / n_kv, b = bg / G, g = bg % G;
    float acc[4] = {bb[0], bb[1], bb[2], bb[3]};
    float xv[kMaxTaps][4];
    const h16* base = q + (size_t)b * n * C + g * Cg + c0;
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) {
      const int p = stride * j - pad + t;
      if (t < ks && p >= 0 && p < n) {
        load4(base + (size_t)p * C, xv[t]);
      } else {
        xv[t][0] = xv[t][1] = xv[t][2] = xv[t][3] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fmaf(w[e][t], xv[t][e], acc[e]);
    }
    float ge[4], u = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ge[e] = gelu_erf(acc[e]);
      u = fmaf(ww2[e], ge[e], u);
    }
    u = warp_sum(u);
    const float th = tanhf(u);
    const float du = d_off[item] * offset_scale * (1.0f - th * th);
    float dyv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gw2[e] = fmaf(du, ge[e], gw2[e]);
      dyv[e] = du * ww2[e] * gelu_erf_grad(acc[e]);
      gb0[e] += dyv[e];
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t) gw0[e][t] = fmaf(dyv[e], xv[t][e], gw0[e][t]);
    }
    *reinterpret_cast<float4*>(dy + (size_t)item * Cg + c0) = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
  }
  // CTA reduction of the parameter gradients
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) red[wib][(c0 + e) * (kMaxTaps + 2) + t] = gw0[e][t];
    red[wib][(c0 + e) * (kMaxTaps + 2) + kMaxTaps] = gb0[e];
    red[wib][(c0 + e) * (kMaxTaps + 2) + kMaxTaps + 1] = gw2[e];
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 128 * (kMaxTaps + 2); i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nw; ++k) s += red[k][i];
    const int c = i / (kMaxTaps + 2), t = i % (kMaxTaps + 2);
    if (t < kMaxTaps) {
      if (t < ks) atomicAdd(wgrad + c * ks + t, s);
    } else if (t == kMaxTaps) {
      atomicAdd(wgrad + Cg * ks + c, s);
    } else {
      atomicAdd(wgrad + Cg * ks + Cg + c, s);
    }
  }
}

// offsets backward, pass B (elementwise over [B, n, C]):
//   dq[b,p,c] = dq_attn[b,p,c] * scale + sum_{j : 0 <= p + pad - stride*j < ks} dy[(b,g), j, c'] * w0[c', p + pad - stride*j]
// writes the total query gradient in fp32 (operand of the TF32 dWq / dx1 GEMMs).
__global__ void __launch_bounds__(256)
offsets_dq_combine_kernel(const float* __restrict__ dq_attn, const float* __restrict__ dy,
                          const float* __restrict__ w0, int B, int n, int C, int G, int ks, int stride, int pad,
                          int n_kv, float scale, float* __restrict__ dq, dml::bf16* __restrict__ dq_pair, long long plane) {
  // C / 4 threads per token (host guarantees that C / 4 divides the block size): the channel of a thread never changes,
  // tokens advance by whole blocks - no per-element divisions (they made the first version compute-bound: 30 us for a
  // 67 MB pass)
  const int Cg = C / G;
  const int tpt = C >> 2;                                  // threads per token
  const int c = (threadIdx.x % tpt) * 4;
  const int g = c / Cg, cc = c % Cg;
  const int tok_per_blk = blockDim.x / tpt;
  const long long rows = (long long)B * n;
  for (long long row = (long long)blockIdx.x * tok_per_blk + threadIdx.x / tpt; row < rows; row += (long long)gridDim.x * tok_per_blk) {
    const int b = (int)(row / n), p = (int)(row - (long long)b * n);
    const size_t i = (size_t)row * C + c;
    const float4 a = *reinterpret_cast<const float4*>(dq_attn + i);
    float r[4] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale};
    const int jhi = min((p + pad) / stride, n_kv - 1);
    for (int j = jhi; j >= 0; --j) {                     // at most ceil(ks / stride) keys see this token
      const int t = p + pad - stride * j;
      if (t >= ks) break;
      const float4 d = *reinterpret_cast<const float4*>(dy + ((size_t)(b * G + g) * n_kv + j) * Cg + cc);
      r[0] = fmaf(d.x, __ldg(w0 + (cc + 0) * ks + t), r[0]);
      r[1] = fmaf(d.y, __ldg(w0 + (cc + 1) * ks + t), r[1]);
      r[2] = fmaf(d.z, __ldg(w0 + (cc + 2) * ks + t), r[2]);
      r[3] = fmaf(d.w, __ldg(w0 + (cc + 3) * ks + t), r[3]);
    }
    if (dq) *reinterpret_cast<float4*>(dq + i) = make_float4(r[0], r[1], r[2], r[3]);
    if (dq_pair) dml::store_pair4(dq_pair + i, dq_pair + plane + i, r[0], r[1], r[2], r[3]);     // operand of the dWq / dx1 GEMMs
  }
}

// ------------------------------------------------------------------------------------------------
// key/value gather.  Shipped semantics (quirk T1): the sampling grid's learned coordinate lands on the
// size-1 W axis, y is always 0 => every key j samples the CENTRE of the sequence (taps i0, i1 with
// weights wy0, wy1 given by the host from iy = (n-1)/2) times the x tent weight 1-|g/2|.
//   kv[b, j, c] = (x2[b,i0,c] wy0 + x2[b,i1,c] wy1) * tent(g[(b,grp(c)), j])       (fp32 out, token-major)
// x2 [B, n, dim] fp32 token-major; dim/G channels per group; one warp per (b, j), lane owns dim/32 channels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tent_weight(float g, float* dtent) {
  // same fp32 operation order as ATen's grid_sampler on a width-1 image (align_corners = False)
  const float ix = ((g + 1.0f) * 1.0f - 1.0f) / 2.0f;
  const float x0 = floorf(ix);
  float w = 0.f, d = 0.f;
  if (x0 == 0.0f) { w = (x0 + 1.0f) - ix; d = -0.5f; }
  else if (x0 == -1.0f) { w = ix - x0; d = 0.5f; }
  if (dtent) *dtent = d;
  return w;
}

template <int VPL>  // channels per lane (dim = 32 * VPL)
__global__ void __launch_bounds__(256)
kv_gather_fwd_kernel(const float* __restrict__ x2, const float* __restrict__ gnorm, int B, int n, int dim, int G,
                     int n_kv, int i0, int i1, float wy0, float wy1, float* __restrict__ kv) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int c0 = lane * VPL;
  const int grp = c0 / (dim / G);
  const int total = B * n_kv;
  int cur_b = -1;
  float cen[VPL];
  for (int item = warp; item < total; item += nwarps) {
    const int b = item / n_kv, j = item % n_kv;
    if (b != cur_b) {
      cur_b = b;
#pragma unroll
      for (int e = 0; e < VPL; ++e) {
        float v = x2[((size_t)b * n + i0) * dim + c0 + e] * wy0;
        if (wy1 != 0.f) v = fmaf(x2[((size_t)b * n + i1) * dim + c0 + e], wy1, v);
        cen[e] = v;
      }
    }
    const float tw = tent_weight(gnorm[(size_t)(b * G + grp) * n_kv + j], nullptr);
    float* o = kv + ((size_t)b * n_kv + j) * dim + c0;
    static_assert(VPL == 4, "one float4 per lane");
    *reinterpret_cast<float4*>(o) = make_float4(cen[0] * tw, cen[1] * tw, cen[2] * tw, cen[3] * tw);
  }
}

// backward: dcentre[b,c] += sum_j dkv[b,j,c] tent(g_j);   dg[(b,grp), j] += sum_{c in grp} dkv[b,j,c] centre[b,c] dtent
// dkv [B, n_kv, dim] fp32.  dcentre [B, dim] and dg must be initialised by the caller (dg is accumulated into).
template <int VPL>
__global__ void __launch_bounds__(256)
kv_gather_bwd_kernel(const float* __restrict__ x2, const float* __restrict__ gnorm, const float* __restrict__ dkv,
                     int B, int n, int dim, int G, int n_kv, int i0, int i1, float wy0, float wy1,
                     float* __restrict__ dcentre, float* __restrict__ dg) {
  __shared__ float red[8][32 * VPL];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = lane * VPL;
  const int cpg = dim / G;             // channels per group
  const int lanes_per_grp = cpg / VPL;  // lanes that share one group (power of two)
  const int grp = c0 / cpg;
  const int b = blockIdx.y;
  float cen[VPL], acc[VPL];
#pragma unroll
  for (int e = 0; e < VPL; ++e) {
    float v = x2[((size_t)b * n + i0) * dim + c0 + e] * wy0;
    if (wy1 != 0.f) v = fmaf(x2[((size_t)b * n + i1) * dim + c0 + e], wy1, v);
    cen[e] = v;
    acc[e] = 0.f;
  }
  const int warp = blockIdx.x * (blockDim.x >> 5) + wib;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int j = warp; j < n_kv; j += nwarps) {
    float dt;
    const size_t gi = (size_t)(b * G + grp) * n_kv + j;
    const float tw = tent_weight(gnorm[gi], &dt);
    const float* d = dkv + ((size_t)b * n_kv + j) * dim + c0;
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < VPL; ++e) {
      const float dv = d[e];
      acc[e] = fmaf(dv, tw, acc[e]);
      s = fmaf(dv, cen[e], s);
    }
    for (int o = lanes_per_grp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((lane & (lanes_per_grp - 1)) == 0) dg[gi] += s * dt;
  }
#pragma unroll
  for (int e = 0; e < VPL; ++e) red[wib][c0 + e] = acc[e];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VPL; i += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k][i];
    atomicAdd(dcentre + (size_t)b * dim + i, s);
  }
}

}  // namespace dml

extern "C" {

int dml_offsets_kv_len(int n, int ksize, int stride) {
  const int pad = (ksize - stride) / 2;
  return (n + 2 * pad - ksize) / stride + 1;
}

int dml_offsets_fwd(const void* q, const float* w0, const float* b0, const float* w2, int B, int n, int C, int G,
                    int ksize, int stride, float offset_scale, float* vgrid, float* gnorm, void* stream) {
  DML_CHECK_ARG(q && w0 && b0 && w2 && vgrid && gnorm && B > 0 && n > 0 && G > 0);
  if (C != G * 128 || ksize > dml::kMaxTaps || ksize < stride || ((ksize - stride) & 1)) return DML_EUNSUPPORTED;
  const int pad = (ksize - stride) / 2, n_kv = dml_offsets_kv_len(n, ksize, stride);
  DML_CHECK_ARG(n_kv >= 1);
  const int warps = B * G * n_kv;
  const int blocks = min(dml::cdiv(warps, 8), 148 * 8);
  dml::offsets_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const dml::h16*)q, w0, b0, w2, B, n, C, G, ksize,
                                                                    stride, pad, n_kv, offset_scale, vgrid, gnorm);
  DML_RETURN_LAUNCH();
}

int dml_offsets_bwd(const void* q, const float* w0, const float* b0, const float* w2, const float* d_off,
                    const float* dq_attn, float attn_scale, int B, int n, int C, int G, int ksize, int stride,
                    float offset_scale, float* dy_ws, float* wgrad, void* dq_out, void* stream) {
  return dml_offsets_bwd_pair(q, w0, b0, w2, d_off, dq_attn, attn_scale, B, n, C, G, ksize, stride, offset_scale, dy_ws, wgrad,
                              dq_out, nullptr, 0, stream);
}

int dml_offsets_bwd_pair(const void* q, const float* w0, const float* b0, const float* w2, const float* d_off,
                         const float* dq_attn, float attn_scale, int B, int n, int C, int G, int ksize, int stride,
                         float offset_scale, float* dy_ws, float* wgrad, void* dq_out, void* dq_pair, long long plane_stride,
                         void* stream) {
  // Two stages, either may be skipped: d_off != NULL runs the offset-network backward (d_off -> dy_ws, wgrad), dq_attn != NULL the
  // combination dq_out / dq_pair = attn_scale * dq_attn + (transposed convolution of dy_ws).  A caller whose dq_attn is still being
  // produced on another stream calls stage one first and stage two once it is there.
  const bool stage1 = d_off != nullptr, stage2 = dq_attn != nullptr;
  DML_CHECK_ARG(q && w0 && b0 && w2 && dy_ws && wgrad && (stage1 || stage2) && B > 0 && n > 0 && G > 0);
  DML_CHECK_ARG(!stage2 || dq_out || dq_pair);
  if (dq_pair && ((((uintptr_t)dq_pair) & 7) || (plane_stride & 3))) return DML_EINVAL;
  if (C != G * 128 || ksize > dml::kMaxTaps || ksize < stride || ((ksize - stride) & 1)) return DML_EUNSUPPORTED;
  const int pad = (ksize - stride) / 2, n_kv = dml_offsets_kv_len(n, ksize, stride);
  const int Cg = C / G;
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 != 0 || 256 % (C / 4) != 0 || ((C / G) % 4) != 0) return DML_EUNSUPPORTED;
  if (stage1) {
    cudaError_t e = cudaMemsetAsync(wgrad, 0, sizeof(float) * (size_t)(Cg * ksize + 2 * Cg), st);
    if (e != cudaSuccess) return (int)e;
    const int warps = B * G * n_kv;
    const int blocks = min(dml::cdiv(warps, 8), 148 * 2);
    dml::offsets_bwd_kernel<<<blocks, 256, 0, st>>>((const dml::h16*)q, w0, b0, w2, d_off, B, n, C, G, ksize, stride, pad,
                                                   n_kv, offset_scale, dy_ws, wgrad);
  }
  if (stage2) {
    const size_t total4 = (size_t)B * n * C / 4;
    const int blocks2 = (int)min((total4 + 255) / 256, (size_t)148 * 16);
    dml::offsets_dq_combine_kernel<<<blocks2, 256, 0, st>>>(dq_attn, dy_ws, w0, B, n, C, G, ksize, stride, pad, n_kv,
                                                           attn_scale, (float*)dq_out, (dml::bf16*)dq_pair, plane_stride);
  }
  DML_RETURN_LAUNCH();
}

int dml_kv_gather_fwd(const float* x2, const float* gnorm, int B, int n, int dim, int G, int n_kv, int i0, int i1,
                      float wy0, float wy1, void* kv, void* stream) {
  DML_CHECK_ARG(x2 && gnorm && kv && B > 0 && n > 0 && n_kv > 0 && i0 >= 0 && i0 < n);
  if (dim != 128 || (dim % G) != 0 || ((dim / G) % 4) != 0) return DML_EUNSUPPORTED;
  if (wy1 != 0.f) DML_CHECK_ARG(i1 >= 0 && i1 < n);
  const int warps = B * n_kv;
  const int blocks = min(dml::cdiv(warps, 8), 148 * 8);
  dml::kv_gather_fwd_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(x2, gnorm, B, n, dim, G, n_kv, i0, i1, wy0, wy1,
                                                                        (float*)kv);
  DML_RETURN_LAUNCH();
}

int dml_kv_gather_bwd(const float* x2, const float* gnorm, const float* dkv, int B, int n, int dim, int G, int n_kv,
                      int i0, int i1, float wy0, float wy1, float* dcentre, float* dg, void* stream) {
  DML_CHECK_ARG(x2 && gnorm && dkv && dcentre && dg && B > 0 && n > 0 && n_kv > 0 && i0 >= 0 && i0 < n);
  if (dim != 128 || (dim % G) != 0 || ((dim / G) % 4) != 0) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dcentre, 0, sizeof(float) * (size_t)B * dim, st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(min(dml::cdiv(n_kv, 8), 64), B);
  dml::kv_gather_bwd_kernel<4><<<grid, 256, 0, st>>>(x2, gnorm, dkv, B, n, dim, G, n_kv, i0, i1, wy0, wy1, dcentre, dg);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
