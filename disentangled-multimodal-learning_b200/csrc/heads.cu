// The small dense heads behind the attention layer, one kernel per direction each.  Between the forward and the backward of
// a 16k-patch bag the GPU has nothing else to run, so every microsecond-sized launch of these heads (LayerNorm of the cls
// row, four tiny Linear layers, their ~40 autograd kernels) is a bubble in a step of under 4 ms.
//   tower head   (models/DeformCrossTransMIL.py:128-151)  h = norm(x)[:, 0];  logits = _fc2(h);  encoded = multimodal_projection(h)
//   linear3 head (models/model.py:535-558)                hazard = classifier(cat(vt, vi)), hazard_tumor = classifier_tumor(vt),
//                                                          hazard_immune = classifier_immune(vi)  [+ sigmoid for survival]
// One CTA per bag row; a warp computes an output unit (lanes over the inputs, shuffle reduction).
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {
namespace hd {

constexpr int kThreads = 512, kMaxD = 512;
// The tower-head kernels are ONE CTA per bag on the critical path between the forward and the backward of the step: both
// weight matrices ([nc + De, D]) are requested into (dynamic) shared memory by all threads at the top - one global-memory
// latency instead of one per output unit - when they fit; otherwise they are read through the read-only cache as before.
constexpr int kStageFloats = 48 * 1024;

__device__ __forceinline__ bool stage_rows(float* stage, const float* __restrict__ W2, int nc, const float* __restrict__ Wp, int De, int D) {
  if ((nc + De) * D > kStageFloats) return false;
  for (int i = threadIdx.x; i < nc * D; i += kThreads) stage[i] = __ldg(W2 + i);
  for (int i = threadIdx.x; i < De * D; i += kThreads) stage[nc * D + i] = __ldg(Wp + i);
  return true;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int k = 0; k < kThreads / 32; ++k) s += red[k];
  return s;
}

// x: row `row` of [B, n, D] (row stride ld = n * D between bags).  hn [B, D] saved (the normalised row), stats [B, 2] = mean, rstd.
__global__ void __launch_bounds__(kThreads)
tower_head_fwd_kernel(const float* __restrict__ x, long long bag_stride, int D, const float* __restrict__ lw, const float* __restrict__ lb,
                      float eps, const float* __restrict__ W2, const float* __restrict__ b2, int nc, const float* __restrict__ Wp,
                      const float* __restrict__ bp, int De, float* __restrict__ hn, float* __restrict__ stats,
                      float* __restrict__ logits, float* __restrict__ enc) {
  extern __shared__ float stage[];
  __shared__ float h[kMaxD], red[kThreads / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  const bool staged = stage_rows(stage, W2, nc, Wp, De, D);
  const float* xr = x + (size_t)b * bag_stride;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) { h[i] = xr[i]; s += h[i]; }
  const float mu = block_sum(s, red) / D;
  float q = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) { const float d = h[i] - mu; q += d * d; }
  const float rs = rsqrtf(block_sum(q, red) / D + eps);
  for (int i = threadIdx.x; i < D; i += kThreads) {
    const float v = fmaf((h[i] - mu) * rs, __ldg(lw + i), __ldg(lb + i));
    h[i] = v;
    hn[(size_t)b * D + i] = v;
  }
  if (threadIdx.x == 0) { stats[2 * b] = mu; stats[2 * b + 1] = rs; }
  __syncthreads();
  for (int o = warp; o < nc + De; o += nw) {
    float acc = 0.f;
    if (staged) {
      const float* w = stage + o * D;
      for (int k = lane; k < D; k += 32) acc = fmaf(w[k], h[k], acc);
    } else {
      const float* w = o < nc ? W2 + (size_t)o * D : Wp + (size_t)(o - nc) * D;
      for (int k = lane; k < D; k += 32) acc = fmaf(__ldg(w + k), h[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (o < nc) logits[(size_t)b * nc + o] = acc + __ldg(b2 + o);
      else enc[(size_t)b * De + o - nc] = acc + __ldg(bp + o - nc);
    }
  }
}

// grads: dparams = dlw [D], dlb [D], dW2 [nc, D], db2 [nc], dWp [De, D], dbp [De] (accumulated with atomics into a zeroed buffer);
// dx: row `row` of the zeroed [B, n, D] gradient.
__global__ void __launch_bounds__(kThreads)
tower_head_bwd_kernel(const float* __restrict__ x, long long bag_stride, int D, const float* __restrict__ lw, const float* __restrict__ W2,
                      int nc, const float* __restrict__ Wp, int De, const float* __restrict__ hn, const float* __restrict__ stats,
                      const float* __restrict__ dlogits, const float* __restrict__ denc, float* __restrict__ dparams,
                      float* __restrict__ dx, long long dx_bag_stride) {
  extern __shared__ float stage[];
  __shared__ float g[kMaxD], go[kMaxD + 16], red[kThreads / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  const bool staged = stage_rows(stage, W2, nc, Wp, De, D);
  float* dlw = dparams;
  float* dlb = dlw + D;
  float* dW2 = dlb + D;
  float* db2 = dW2 + (size_t)nc * D;
  float* dWp = db2 + nc;
  float* dbp = dWp + (size_t)De * D;
  for (int o = threadIdx.x; o < nc + De; o += kThreads)
    go[o] = o < nc ? (dlogits ? dlogits[(size_t)b * nc + o] : 0.f) : (denc ? denc[(size_t)b * De + o - nc] : 0.f);
  __syncthreads();
  const float* hb = hn + (size_t)b * D;
  // weight / bias gradients of the two linears
  for (int idx = threadIdx.x; idx < (nc + De) * D; idx += kThreads) {
    const int o = idx / D, k = idx - o * D;
    const float v = go[o] * hb[k];
    if (v != 0.f) atomicAdd((o < nc ? dW2 + (size_t)o * D : dWp + (size_t)(o - nc) * D) + k, v);
  }
  for (int o = threadIdx.x; o < nc + De; o += kThreads)
    if (go[o] != 0.f) atomicAdd(o < nc ? db2 + o : dbp + o - nc, go[o]);
  // d hn[k] = sum_o W[o, k] go[o]
  for (int k0 = warp * 32; k0 < D; k0 += nw * 32) {
    const int k = k0 + lane;
    if (k < D) {
      float s = 0.f;
      if (staged) {
        for (int o = 0; o < nc + De; ++o) s = fmaf(stage[o * D + k], go[o], s);
      } else {
        for (int o = 0; o < nc; ++o) s = fmaf(__ldg(W2 + (size_t)o * D + k), go[o], s);
        for (int o = 0; o < De; ++o) s = fmaf(__ldg(Wp + (size_t)o * D + k), go[nc + o], s);
      }
      g[k] = s;
    }
  }
  __syncthreads();
  // LayerNorm backward on the row
  const float mu = stats[2 * b], rs = stats[2 * b + 1];
  const float* xr = x + (size_t)b * bag_stride;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    const float xh = (xr[i] - mu) * rs, gw = g[i] * __ldg(lw + i);
    s1 += gw;
    s2 += gw * xh;
    atomicAdd(dlw + i, g[i] * xh);
    atomicAdd(dlb + i, g[i]);
  }
  const float m1 = block_sum(s1, red) / D;
  const float m2 = block_sum(s2, red) / D;
  float* dxr = dx + (size_t)b * dx_bag_stride;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    const float xh = (xr[i] - mu) * rs, gw = g[i] * __ldg(lw + i);
    dxr[i] = rs * (gw - m1 - xh * m2);
  }
}

// y_c = act(Wc [nc, Da + Db] cat(a, b) + bc), y_a = act(Wa [nc, Da] a + ba), y_b = act(Wb [nc, Db] b + bb);  act = sigmoid or identity
__global__ void __launch_bounds__(kThreads)
linear3_fwd_kernel(const float* __restrict__ a, const float* __restrict__ bvec, int Da, int Db, const float* __restrict__ Wc,
                   const float* __restrict__ bc, const float* __restrict__ Wa, const float* __restrict__ ba, const float* __restrict__ Wb,
                   const float* __restrict__ bb, int nc, int sigmoid, float* __restrict__ yc, float* __restrict__ ya, float* __restrict__ yb) {
  __shared__ float f[2 * kMaxD];
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  for (int i = threadIdx.x; i < Da; i += kThreads) f[i] = a[(size_t)r * Da + i];
  for (int i = threadIdx.x; i < Db; i += kThreads) f[Da + i] = bvec[(size_t)r * Db + i];
  __syncthreads();
  for (int o = warp; o < 3 * nc; o += nw) {
    const int which = o / nc, u = o - which * nc;
    const float* w = which == 0 ? Wc + (size_t)u * (Da + Db) : (which == 1 ? Wa + (size_t)u * Da : Wb + (size_t)u * Db);
    const float* in = which == 2 ? f + Da : f;
    const int K = which == 0 ? Da + Db : (which == 1 ? Da : Db);
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(w + k), in[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += which == 0 ? __ldg(bc + u) : (which == 1 ? __ldg(ba + u) : __ldg(bb + u));
      if (sigmoid) acc = 1.0f / (1.0f + __expf(-acc));
      (which == 0 ? yc : (which == 1 ? ya : yb))[(size_t)r * nc + u] = acc;
    }
  }
}

// dparams = dWc [nc, Da + Db], dbc [nc], dWa [nc, Da], dba [nc], dWb [nc, Db], dbb [nc] (atomics into a zeroed buffer); da [B, Da], db_ [B, Db]
__global__ void __launch_bounds__(kThreads)
linear3_bwd_kernel(const float* __restrict__ a, const float* __restrict__ bvec, int Da, int Db, const float* __restrict__ Wc,
                   const float* __restrict__ Wa, const float* __restrict__ Wb, int nc, int sigmoid, const float* __restrict__ yc,
                   const float* __restrict__ ya, const float* __restrict__ yb, const float* __restrict__ gyc, const float* __restrict__ gya,
                   const float* __restrict__ gyb, float* __restrict__ dparams, float* __restrict__ da, float* __restrict__ db_) {
  __shared__ float f[2 * kMaxD], g[3 * 64];
  const int r = blockIdx.x;
  const int Dc = Da + Db;
  for (int i = threadIdx.x; i < Da; i += kThreads) f[i] = a[(size_t)r * Da + i];
  for (int i = threadIdx.x; i < Db; i += kThreads) f[Da + i] = bvec[(size_t)r * Db + i];
  for (int o = threadIdx.x; o < 3 * nc; o += kThreads) {
    const int which = o / nc, u = o - which * nc;
    const float* gy = which == 0 ? gyc : (which == 1 ? gya : gyb);
    float v = gy ? gy[(size_t)r * nc + u] : 0.f;
    if (sigmoid) {
      const float y = (which == 0 ? yc : (which == 1 ? ya : yb))[(size_t)r * nc + u];
      v *= y * (1.0f - y);
    }
    g[o] = v;
  }
  __syncthreads();
  float* dWc = dparams;
  float* dbc = dWc + (size_t)nc * Dc;
  float* dWa = dbc + nc;
  float* dba = dWa + (size_t)nc * Da;
  float* dWb = dba + nc;
  float* dbb = dWb + (size_t)nc * Db;
  for (int idx = threadIdx.x; idx < nc * Dc; idx += kThreads) atomicAdd(dWc + idx, g[idx / Dc] * f[idx % Dc]);
  for (int idx = threadIdx.x; idx < nc * Da; idx += kThreads) atomicAdd(dWa + idx, g[nc + idx / Da] * f[idx % Da]);
  for (int idx = threadIdx.x; idx < nc * Db; idx += kThreads) atomicAdd(dWb + idx, g[2 * nc + idx / Db] * f[Da + idx % Db]);
  for (int u = threadIdx.x; u < nc; u += kThreads) { atomicAdd(dbc + u, g[u]); atomicAdd(dba + u, g[nc + u]); atomicAdd(dbb + u, g[2 * nc + u]); }
  for (int k = threadIdx.x; k < Dc; k += kThreads) {
    float s = 0.f;
    for (int u = 0; u < nc; ++u) s = fmaf(__ldg(Wc + (size_t)u * Dc + k), g[u], s);
    if (k < Da) {
      for (int u = 0; u < nc; ++u) s = fmaf(__ldg(Wa + (size_t)u * Da + k), g[nc + u], s);
      da[(size_t)r * Da + k] = s;
    } else {
      for (int u = 0; u < nc; ++u) s = fmaf(__ldg(Wb + (size_t)u * Db + k - Da), g[2 * nc + u], s);
      db_[(size_t)r * Db + k - Da] = s;
    }
  }
}

}  // namespace hd
}  // namespace dml

extern "C" {

int dml_tower_head_fwd(const float* x, long long bag_stride, int B, int D, const float* ln_w, const float* ln_b, float eps,
                       const float* W2, const float* b2, int nc, const float* Wp, const float* bp, int De, float* hn, float* stats,
                       float* logits, float* enc, void* stream) {
  DML_CHECK_ARG(x && ln_w && ln_b && W2 && b2 && Wp && bp && hn && stats && logits && enc && B > 0 && nc > 0 && De > 0);
  if (D <= 0 || D > dml::hd::kMaxD) return DML_EUNSUPPORTED;
  const int smem = dml::hd::kStageFloats * (int)sizeof(float);
  cudaError_t ea = cudaFuncSetAttribute(dml::hd::tower_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) return (int)ea;
  dml::hd::tower_head_fwd_kernel<<<B, dml::hd::kThreads, smem, (cudaStream_t)stream>>>(x, bag_stride, D, ln_w, ln_b, eps, W2, b2, nc, Wp, bp,
                                                                                  De, hn, stats, logits, enc);
  DML_RETURN_LAUNCH();
}

int dml_tower_head_bwd(const float* x, long long bag_stride, int B, int D, const float* ln_w, const float* W2, int nc, const float* Wp,
                       int De, const float* hn, const float* stats, const float* dlogits, const float* denc, float* dparams,
                       float* dx, long long dx_bag_stride, void* stream) {
  DML_CHECK_ARG(x && ln_w && W2 && Wp && hn && stats && dparams && dx && B > 0 && nc > 0 && De > 0);
  if (D <= 0 || D > dml::hd::kMaxD || nc + De > dml::hd::kMaxD + 16) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t np = (size_t)2 * D + (size_t)nc * D + nc + (size_t)De * D + De;
  cudaError_t e = cudaMemsetAsync(dparams, 0, np * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int smem = dml::hd::kStageFloats * (int)sizeof(float);
  e = cudaFuncSetAttribute(dml::hd::tower_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  dml::hd::tower_head_bwd_kernel<<<B, dml::hd::kThreads, smem, st>>>(x, bag_stride, D, ln_w, W2, nc, Wp, De, hn, stats, dlogits, denc, dparams,
                                                                 dx, dx_bag_stride);
  DML_RETURN_LAUNCH();
}

int dml_linear3_fwd(const float* a, const float* b, int B, int Da, int Db, const float* Wc, const float* bc, const float* Wa,
                    const float* ba, const float* Wb, const float* bb, int nc, int sigmoid, float* yc, float* ya, float* yb,
                    void* stream) {
  DML_CHECK_ARG(a && b && Wc && bc && Wa && ba && Wb && bb && yc && ya && yb && B > 0 && nc > 0);
  if (Da <= 0 || Db <= 0 || Da > dml::hd::kMaxD || Db > dml::hd::kMaxD || nc > 64) return DML_EUNSUPPORTED;
  dml::hd::linear3_fwd_kernel<<<B, dml::hd::kThreads, 0, (cudaStream_t)stream>>>(a, b, Da, Db, Wc, bc, Wa, ba, Wb, bb, nc, sigmoid, yc, ya, yb);
  DML_RETURN_LAUNCH();
}

int dml_linear3_bwd(const float* a, const float* b, int B, int Da, int Db, const float* Wc, const float* Wa, const float* Wb, int nc,
                    int sigmoid, const float* yc, const float* ya, const float* yb, const float* gyc, const float* gya,
                    const float* gyb, float* dparams, float* da, float* db, void* stream) {
  DML_CHECK_ARG(a && b && Wc && Wa && Wb && yc && ya && yb && dparams && da && db && B > 0 && nc > 0);
  if (Da <= 0 || Db <= 0 || Da > dml::hd::kMaxD || Db > dml::hd::kMaxD || nc > 64) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t np = (size_t)nc * (Da + Db) + nc + (size_t)nc * Da + nc + (size_t)nc * Db + nc;
  cudaError_t e = cudaMemsetAsync(dparams, 0, np * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  dml::hd::linear3_bwd_kernel<<<B, dml::hd::kThreads, 0, st>>>(a, b, Da, Db, Wc, Wa, Wb, nc, sigmoid, yc, ya, yb, gyc, gya, gyb, dparams, da, db);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
