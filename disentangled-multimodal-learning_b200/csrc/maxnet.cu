// The omic self-normalising MLP in front of each DeformCrossTransMIL tower - MaxNet (models/model.py:173-218): four
// Linear -> ELU -> AlphaDropout blocks (input -> 64 -> 48 -> 32 -> omic_dim) and a final ReLU - as ONE kernel per direction.
// As separate torch ops it is ~40 microsecond-sized launches forward and ~60 backward per tower; the forward chain sits in
// front of FusionNet (DeformCrossTransMIL.py:105-111), the backward chain at the very end of the step, both on the
// critical path of a 16k-patch bag whose whole step is under 4 ms.
// One CTA per bag row; a warp computes an output unit (lanes stride the input, shuffle reduction: coalesced weight reads).
// AlphaDropout (training) follows torch's formula exactly - kept: a x + alpha a p, dropped: alpha a (p - 1), a = ((alpha^2 p + 1)
// (1 - p))^-1/2, alpha = 1.7580993408473766 - with the uniform numbers supplied by the caller (one torch RNG launch).
#include <math.h>

#include "../../include/dml_b200.h"
#include "common.cuh"

namespace dml {
namespace mx {

constexpr int kLayers = 4, kMaxDim = 512, kThreads = 512;
constexpr double kAlpha = 1.7580993408473766;
// Dynamic shared memory: the weight matrices of as many layers as fit (in layer order), requested by every thread at the
// top of the kernel - one global-memory latency for the whole net instead of one per output unit (with one bag per step
// the kernel is ONE CTA, and it sits on the critical path of the step: in front of FusionNet forward, last in the backward).
constexpr int kStageFloats = 48 * 1024;           // 192 KB

struct Net {
  const float* W[kLayers];
  const float* b[kLayers];
  int dim[kLayers + 1];      // dim[0] = input, dim[l + 1] = outputs of layer l
};

// Record layouts (per bag row b):  hsave [B][dim0 + dim1 + dim2 + dim3] = the INPUT of every layer (x, then the post-dropout
// activations);  act, u [B][dim1 + dim2 + dim3 + dim4] = the ELU output of every layer (before dropout) / its uniform number;
// feat [B][dim4] = relu(last post-dropout activation).
__device__ __forceinline__ int in_off(const Net& n, int l) { int o = 0; for (int q = 0; q < l; ++q) o += n.dim[q]; return o; }
__device__ __forceinline__ int out_off(const Net& n, int l) { int o = 0; for (int q = 0; q < l; ++q) o += n.dim[q + 1]; return o; }

// Copies the weights of the leading layers that fit into `stage`; woff[l] = offset of layer l in it, or -1 (read from global).
__device__ __forceinline__ void stage_weights(const Net& net, float* stage, int* woff) {
  int used = 0;
#pragma unroll
  for (int l = 0; l < kLayers; ++l) {
    const int cnt = net.dim[l] * net.dim[l + 1];
    if (used + cnt <= kStageFloats) {
      woff[l] = used;
      for (int i = threadIdx.x; i < cnt; i += kThreads) stage[used + i] = __ldg(net.W[l] + i);
      used += cnt;
    } else {
      woff[l] = -1;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
maxnet_fwd_kernel(const float* __restrict__ x, const Net net, const float* __restrict__ u, float p, float* __restrict__ act,
                  float* __restrict__ hsave, float* __restrict__ feat) {
  extern __shared__ float stage[];
  __shared__ float cur[kMaxDim], nxt[kMaxDim], us[kLayers * kMaxDim], bs[kLayers * kMaxDim];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  const int hstride = in_off(net, kLayers), astride = out_off(net, kLayers);
  int woff[kLayers];
  stage_weights(net, stage, woff);
  for (int i = threadIdx.x; i < net.dim[0]; i += kThreads) cur[i] = x[(size_t)b * net.dim[0] + i];
#pragma unroll
  for (int l = 0; l < kLayers; ++l) {
    const int ao = out_off(net, l);
    for (int i = threadIdx.x; i < net.dim[l + 1]; i += kThreads) {
      bs[ao + i] = __ldg(net.b[l] + i);
      if (u) us[ao + i] = __ldg(u + (size_t)b * astride + ao + i);
    }
  }
  __syncthreads();
  const float a = (float)(1.0 / sqrt((kAlpha * kAlpha * (double)p + 1.0) * (1.0 - (double)p)));
  const float keep_add = (float)kAlpha * a * p, drop_val = (float)kAlpha * a * (p - 1.0f);
#pragma unroll
  for (int l = 0; l < kLayers; ++l) {
    const int din = net.dim[l], dout = net.dim[l + 1], ho = in_off(net, l), ao = out_off(net, l);
    for (int i = threadIdx.x; i < din; i += kThreads) hsave[(size_t)b * hstride + ho + i] = cur[i];
    for (int o = warp; o < dout; o += nw) {
      float s = 0.f;
      if (woff[l] >= 0) {
        const float* w = stage + woff[l] + o * din;
        for (int k = lane; k < din; k += 32) s = fmaf(w[k], cur[k], s);
      } else {
        const float* w = net.W[l] + (size_t)o * din;
        for (int k = lane; k < din; k += 32) s = fmaf(__ldg(w + k), cur[k], s);
      }
      s = warp_sum(s);
      if (lane == 0) {
        s += bs[ao + o];
        const float y = s > 0.f ? s : expm1f(s);                       // ELU
        act[(size_t)b * astride + ao + o] = y;
        float h = y;
        if (u) h = us[ao + o] < p ? drop_val : fmaf(a, y, keep_add);      // AlphaDropout
        nxt[o] = h;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < dout; i += kThreads) cur[i] = nxt[i];
    __syncthreads();
  }
  const int dl = net.dim[kLayers];
  for (int i = threadIdx.x; i < dl; i += kThreads) feat[(size_t)b * dl + i] = fmaxf(cur[i], 0.f);
}

// backward: dfeat [B][dim4] -> dparams (ACCUMULATED with atomics into a buffer the entry point zeroes; layout: for each layer
// dW [dout][din] then db [dout]) and dx [B][dim0] (may be NULL).
__global__ void __launch_bounds__(kThreads)
maxnet_bwd_kernel(const float* __restrict__ dfeat, const Net net, const float* __restrict__ u, float p, const float* __restrict__ act,
                  const float* __restrict__ hsave, const float* __restrict__ feat, float* __restrict__ dparams, float* __restrict__ dx) {
  extern __shared__ float stage[];
  __shared__ float delta[kMaxDim], dprev[kMaxDim], hs[kLayers * kMaxDim], as[kLayers * kMaxDim], us[kLayers * kMaxDim];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads / 32;
  const int hstride = in_off(net, kLayers), astride = out_off(net, kLayers);
  const float a = (float)(1.0 / sqrt((kAlpha * kAlpha * (double)p + 1.0) * (1.0 - (double)p)));
  int poff[kLayers];
  {
    int o = 0;
#pragma unroll
    for (int l = 0; l < kLayers; ++l) { poff[l] = o; o += net.dim[l + 1] * net.dim[l] + net.dim[l + 1]; }
  }
  int woff[kLayers];
  stage_weights(net, stage, woff);
  for (int i = threadIdx.x; i < hstride; i += kThreads) hs[i] = hsave[(size_t)b * hstride + i];
  for (int i = threadIdx.x; i < astride; i += kThreads) {
    as[i] = act[(size_t)b * astride + i];
    if (u) us[i] = __ldg(u + (size_t)b * astride + i);
  }
  const int dl = net.dim[kLayers];
  for (int i = threadIdx.x; i < dl; i += kThreads)          // through the final ReLU
    dprev[i] = feat[(size_t)b * dl + i] > 0.f ? dfeat[(size_t)b * dl + i] : 0.f;
  __syncthreads();
#pragma unroll
  for (int l = kLayers - 1; l >= 0; --l) {
    const int din = net.dim[l], dout = net.dim[l + 1], ho = in_off(net, l), ao = out_off(net, l);
    // delta = gradient of the pre-activation: AlphaDropout (kept units scale by a, dropped units pass nothing), ELU' = y > 0 ? 1 : y + 1
    for (int o = threadIdx.x; o < dout; o += kThreads) {
      float g = dprev[o];
      if (u) g = us[ao + o] < p ? 0.f : g * a;
      const float y = as[ao + o];
      delta[o] = g * (y > 0.f ? 1.0f : y + 1.0f);
    }
    __syncthreads();
    const float* hin = hs + ho;
    float* dW = dparams + poff[l];
    float* db = dW + dout * din;
    for (int idx = threadIdx.x; idx < dout * din; idx += kThreads) {
      const int o = idx / din, k = idx - o * din;
      atomicAdd(dW + idx, delta[o] * hin[k]);
    }
    for (int o = threadIdx.x; o < dout; o += kThreads) atomicAdd(db + o, delta[o]);
    if (l > 0 || dx) {
      // d(input)[k] = sum_o W[o, k] delta[o]: lanes over k (consecutive weight columns)
      for (int k0 = warp * 32; k0 < din; k0 += nw * 32) {
        const int k = k0 + lane;
        if (k < din) {
          float s = 0.f;
          if (woff[l] >= 0) {
            const float* w = stage + woff[l] + k;
            for (int o = 0; o < dout; ++o) s = fmaf(w[o * din], delta[o], s);
          } else {
            for (int o = 0; o < dout; ++o) s = fmaf(__ldg(net.W[l] + (size_t)o * din + k), delta[o], s);
          }
          if (l > 0) dprev[k] = s;                            // dprev is no longer read in this layer (delta is complete)
          else dx[(size_t)b * din + k] = s;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace mx
}  // namespace dml

extern "C" {

static int mx_fill(dml::mx::Net& net, const float* const* W, const float* const* b, const int* dims) {
  for (int l = 0; l <= dml::mx::kLayers; ++l) {
    if (dims[l] <= 0 || dims[l] > dml::mx::kMaxDim) return DML_EUNSUPPORTED;
    net.dim[l] = dims[l];
  }
  for (int l = 0; l < dml::mx::kLayers; ++l) {
    if (!W[l] || !b[l]) return DML_EINVAL;
    net.W[l] = W[l];
    net.b[l] = b[l];
  }
  return DML_OK;
}

int dml_maxnet_fwd(const float* x, const float* const* W, const float* const* b, const int* dims, int B, const float* u, float p,
                   float* act, float* hsave, float* feat, void* stream) {
  DML_CHECK_ARG(x && W && b && dims && B > 0 && act && hsave && feat && p >= 0.f && p < 1.f);
  dml::mx::Net net;
  int rc = mx_fill(net, W, b, dims);
  if (rc) return rc;
  const int smem = dml::mx::kStageFloats * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(dml::mx::maxnet_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  dml::mx::maxnet_fwd_kernel<<<B, dml::mx::kThreads, smem, (cudaStream_t)stream>>>(x, net, u, p, act, hsave, feat);
  DML_RETURN_LAUNCH();
}

int dml_maxnet_bwd(const float* dfeat, const float* const* W, const float* const* b, const int* dims, int B, const float* u, float p,
                   const float* act, const float* hsave, const float* feat, float* dparams, float* dx, void* stream) {
  DML_CHECK_ARG(dfeat && W && b && dims && B > 0 && act && hsave && feat && dparams);
  dml::mx::Net net;
  int rc = mx_fill(net, W, b, dims);
  if (rc) return rc;
  size_t n = 0;
  for (int l = 0; l < dml::mx::kLayers; ++l) n += (size_t)dims[l + 1] * dims[l] + dims[l + 1];
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dparams, 0, n * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int smem = dml::mx::kStageFloats * (int)sizeof(float);
  e = cudaFuncSetAttribute(dml::mx::maxnet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  dml::mx::maxnet_bwd_kernel<<<B, dml::mx::kThreads, smem, st>>>(dfeat, net, u, p, act, hsave, feat, dparams, dx);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
