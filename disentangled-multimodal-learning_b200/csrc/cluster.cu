// ClusterMergeNet (models/ClusterMergeNet.py:68-207): DPC-KNN clustering and the weighted token merge.
// The reference materialises the N x N distance matrix three times (cdist, the masked copy, the gathered rows: 40 GB each at
// N = 99 856); here the distances are recomputed tile by tile in shared memory and only O(N) results leave the SM:
//   dpc_density : per token the 5 smallest distances (self included) -> exp(-mean d^2) + noise, and the row maximum of d^2
//   dpc_parent  : per token the distance to the nearest token of higher density (or dist_max)
//   dpc_assign  : per token the nearest of the K selected centres
//   merge_fwd / merge_bwd : weighted mean of the tokens of a cluster and its adjoint.
// x: fp32 [B, N, 128] (the LayerNorm output); distances are sums of squared differences in fp32 divided by C.
#include "common.cuh"

namespace dml {
namespace {

constexpr int kCc = 128;       // channels
constexpr int kTR = 32;        // rows (tokens) per CTA
constexpr int kTJ = 32;        // columns per tile
constexpr int kLd = 132;       // shared row stride (floats): 16-byte aligned, rows 8 apart land in distinct banks

struct Tile {
  float xi[kTR * kLd];
  float xj[kTJ * kLd];
};

__device__ __forceinline__ void load_rows(float* dst, const float* __restrict__ x, int b, int N, int r0) {
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int r = i >> 5, c4 = i & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < N) v = *reinterpret_cast<const float4*>(x + ((size_t)b * N + r0 + r) * kCc + c4 * 4);
    *reinterpret_cast<float4*>(dst + r * kLd + c4 * 4) = v;
  }
}

// squared distances of row r to columns cs, cs + 8, cs + 16, cs + 24 of the tile
__device__ __forceinline__ void tile_d2(const Tile& T, int r, int cs, float (&d2)[4]) {
  d2[0] = d2[1] = d2[2] = d2[3] = 0.f;
  const float* xi = T.xi + r * kLd;
#pragma unroll 4
  for (int c = 0; c < kCc; c += 4) {
    const float4 a = *reinterpret_cast<const float4*>(xi + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 b = *reinterpret_cast<const float4*>(T.xj + (cs + 8 * u) * kLd + c);
      const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
      d2[u] = fmaf(e0, e0, d2[u]);
      d2[u] = fmaf(e1, e1, d2[u]);
      d2[u] = fmaf(e2, e2, d2[u]);
      d2[u] = fmaf(e3, e3, d2[u]);
    }
  }
}

__device__ __forceinline__ void insert5(float (&best)[5], float v) {
  if (v < best[4]) {
    best[4] = v;
#pragma unroll
    for (int k = 4; k > 0; --k)
      if (best[k] < best[k - 1]) {
        const float t = best[k];
        best[k] = best[k - 1];
        best[k - 1] = t;
      }
  }
}

// density[b, i] = exp(-mean_{5 nearest} d^2 / C) + 1e-6 noise[b, i];  rowmax2[b, i] = max_j d^2(i, j)
__global__ void __launch_bounds__(256) dpc_density_kernel(const float* __restrict__ x, const float* __restrict__ noise, int N,
                                                          float* __restrict__ density, float* __restrict__ rowmax2) {
  __shared__ __align__(16) Tile T;
  __shared__ float cand[kTR][8][6];
  const int b = blockIdx.y, r0 = blockIdx.x * kTR;
  const int r = threadIdx.x >> 3, cs = threadIdx.x & 7;
  load_rows(T.xi, x, b, N, r0);
  float best[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
  float mx = 0.f;
  for (int j0 = 0; j0 < N; j0 += kTJ) {
    __syncthreads();
    load_rows(T.xj, x, b, N, j0);
    __syncthreads();
    float d2[4];
    tile_d2(T, r, cs, d2);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + cs + 8 * u < N) {
        insert5(best, d2[u]);
        mx = fmaxf(mx, d2[u]);
      }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) cand[r][cs][k] = best[k];
  cand[r][cs][5] = mx;
  __syncthreads();
  if (cs == 0 && r0 + r < N) {
    float b5[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    float m2 = 0.f;
    for (int s = 0; s < 8; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) insert5(b5, cand[r][s][k]);
      m2 = fmaxf(m2, cand[r][s][5]);
    }
    // reference: dist = cdist / sqrt(C); density = exp(-mean(dist_nearest^2)) + rand * 1e-6   (ClusterMergeNet.py:88-104)
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float d = sqrtf(b5[k]) / sqrtf((float)kCc);
      s += d * d;
    }
    const size_t at = (size_t)b * N + r0 + r;
    density[at] = expf(-(s / 5.f)) + noise[at] * 1e-6f;
    rowmax2[at] = m2;
  }
}

// parent[b, i] = min(dist_max[b], min_{j: density_j > density_i} d(i, j)),  d = sqrt(d2) / sqrt(C)   (:111-114)
__global__ void __launch_bounds__(256) dpc_parent_kernel(const float* __restrict__ x, const float* __restrict__ density,
                                                         const float* __restrict__ dist_max, int N, float* __restrict__ parent) {
  __shared__ __align__(16) Tile T;
  __shared__ float dj[kTJ];
  __shared__ float cand[kTR][8];
  const int b = blockIdx.y, r0 = blockIdx.x * kTR;
  const int r = threadIdx.x >> 3, cs = threadIdx.x & 7;
  load_rows(T.xi, x, b, N, r0);
  const float di = (r0 + r < N) ? density[(size_t)b * N + r0 + r] : INFINITY;
  float best = INFINITY;
  for (int j0 = 0; j0 < N; j0 += kTJ) {
    __syncthreads();
    load_rows(T.xj, x, b, N, j0);
    if (threadIdx.x < kTJ) dj[threadIdx.x] = (j0 + threadIdx.x < N) ? density[(size_t)b * N + j0 + threadIdx.x] : -INFINITY;
    __syncthreads();
    float d2[4];
    tile_d2(T, r, cs, d2);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dj[cs + 8 * u] > di) best = fminf(best, d2[u]);
  }
  cand[r][cs] = best;
  __syncthreads();
  if (cs == 0 && r0 + r < N) {
    float m2 = INFINITY;
    for (int s = 0; s < 8; ++s) m2 = fminf(m2, cand[r][s]);
    const float dm = dist_max[b];
    parent[(size_t)b * N + r0 + r] = (m2 == INFINITY) ? dm : fminf(dm, sqrtf(m2) / sqrtf((float)kCc));
  }
}

// idx[b, i] = argmin_c d(x[b, centres[b, c]], x[b, i]) (first minimum), one warp per token   (:121-123)
__global__ void __launch_bounds__(256) dpc_assign_kernel(const float* __restrict__ x, const long long* __restrict__ centres, int N, int K,
                                                         long long* __restrict__ idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
  if (i >= N) return;
  const float4 xi = *reinterpret_cast<const float4*>(x + ((size_t)b * N + i) * kCc + lane * 4);
  float best = INFINITY;
  int arg = 0;
  for (int c = 0; c < K; ++c) {
    const long long ci = centres[(size_t)b * K + c];
    const float4 xc = *reinterpret_cast<const float4*>(x + ((size_t)b * N + ci) * kCc + lane * 4);
    const float e0 = xi.x - xc.x, e1 = xi.y - xc.y, e2 = xi.z - xc.z, e3 = xi.w - xc.w;
    float d = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, e3 * e3)));
    d = warp_sum(d);
    if (d < best) {
      best = d;
      arg = c;
    }
  }
  if (lane == 0) idx[(size_t)b * N + i] = arg;
}

// merged[b, c, :] = sum_{i in c} x_i w_i / W_c,  W_c = sum_{i in c} w_i + 1e-6   (:151-166); one CTA per (c, b), 128 threads
__global__ void __launch_bounds__(128) merge_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const long long* __restrict__ idx, int N, int K, float* __restrict__ merged,
                                                        float* __restrict__ all_w) {
  const int c = blockIdx.x, b = blockIdx.y, k = threadIdx.x;
  const long long* ib = idx + (size_t)b * N;
  const float* wb = w + (size_t)b * N;
  float W = 0.f;
  for (int i = 0; i < N; ++i)
    if (ib[i] == c) W += wb[i];
  W += 1e-6f;
  float acc = 0.f;
  for (int i = 0; i < N; ++i)
    if (ib[i] == c) acc += x[((size_t)b * N + i) * kCc + k] * (wb[i] / W);
  merged[((size_t)b * K + c) * kCc + k] = acc;
  if (k == 0) all_w[(size_t)b * K + c] = W;
}

// dx_i = dmerged[c_i] w_i / W,  dw_i = dmerged[c_i] . (x_i - merged[c_i]) / W; one warp per token
__global__ void __launch_bounds__(256) merge_bwd_kernel(const float* __restrict__ dmerged, const float* __restrict__ x,
                                                        const float* __restrict__ w, const long long* __restrict__ idx,
                                                        const float* __restrict__ merged, const float* __restrict__ all_w, int N, int K,
                                                        float* __restrict__ dx, float* __restrict__ dw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
  if (i >= N) return;
  const size_t at = (size_t)b * N + i;
  const int c = (int)idx[at];
  const float W = all_w[(size_t)b * K + c], nw = w[at] / W;
  const float4 g = *reinterpret_cast<const float4*>(dmerged + ((size_t)b * K + c) * kCc + lane * 4);
  const float4 xv = *reinterpret_cast<const float4*>(x + at * kCc + lane * 4);
  const float4 mv = *reinterpret_cast<const float4*>(merged + ((size_t)b * K + c) * kCc + lane * 4);
  *reinterpret_cast<float4*>(dx + at * kCc + lane * 4) = make_float4(g.x * nw, g.y * nw, g.z * nw, g.w * nw);
  float d = g.x * (xv.x - mv.x) + g.y * (xv.y - mv.y) + g.z * (xv.z - mv.z) + g.w * (xv.w - mv.w);
  d = warp_sum(d);
  if (lane == 0) dw[at] = d / W;
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {

int dml_dpc_density(const float* x, const float* noise, int B, int N, int C, float* density, float* rowmax2, void* stream) {
  DML_CHECK_ARG(x && noise && density && rowmax2 && B > 0 && N >= 5 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  dpc_density_kernel<<<dim3(cdiv(N, kTR), B), 256, 0, (cudaStream_t)stream>>>(x, noise, N, density, rowmax2);
  DML_RETURN_LAUNCH();
}

int dml_dpc_parent(const float* x, const float* density, const float* dist_max, int B, int N, int C, float* parent, void* stream) {
  DML_CHECK_ARG(x && density && dist_max && parent && B > 0 && N > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  dpc_parent_kernel<<<dim3(cdiv(N, kTR), B), 256, 0, (cudaStream_t)stream>>>(x, density, dist_max, N, parent);
  DML_RETURN_LAUNCH();
}

int dml_dpc_assign(const float* x, const long long* centres, int B, int N, int C, int K, long long* idx, void* stream) {
  DML_CHECK_ARG(x && centres && idx && B > 0 && N > 0 && K > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  dpc_assign_kernel<<<dim3(cdiv(N, 8), B), 256, 0, (cudaStream_t)stream>>>(x, centres, N, K, idx);
  DML_RETURN_LAUNCH();
}

int dml_merge_fwd(const float* x, const float* w, const long long* idx, int B, int N, int C, int K, float* merged, float* all_w,
                  void* stream) {
  DML_CHECK_ARG(x && w && idx && merged && all_w && B > 0 && N > 0 && K > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  merge_fwd_kernel<<<dim3(K, B), 128, 0, (cudaStream_t)stream>>>(x, w, idx, N, K, merged, all_w);
  DML_RETURN_LAUNCH();
}

int dml_merge_bwd(const float* dmerged, const float* x, const float* w, const long long* idx, const float* merged, const float* all_w,
                  int B, int N, int C, int K, float* dx, float* dw, void* stream) {
  DML_CHECK_ARG(dmerged && x && w && idx && merged && all_w && dx && dw && B > 0 && N > 0 && K > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  merge_bwd_kernel<<<dim3(cdiv(N, 8), B), 256, 0, (cudaStream_t)stream>>>(dmerged, x, w, idx, merged, all_w, N, K, dx, dw);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
