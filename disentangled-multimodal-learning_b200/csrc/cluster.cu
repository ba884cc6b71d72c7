// ClusterMergeNet (models/ClusterMergeNet.py:68-207): DPC-KNN clustering and the weighted token merge.
// The reference materialises the N x N distance matrix three times (cdist, the masked copy, the gathered rows: 40 GB each at
// N = 99 856); here the distances are recomputed tile by tile in shared memory and only O(N) results leave the SM:
//   dpc_split   : x -> scaled fp16 pair (22 bits) + |x|^2
//   dpc_density : per token the 5 smallest distances (self included) -> exp(-mean d^2) + noise, and the row maximum of d^2
//   dpc_parent  : per token the distance to the nearest token of higher density (or dist_max)
//   dpc_assign  : per token the nearest of the K selected centres
//   merge_fwd / merge_bwd : weighted mean of the tokens of a cluster and its adjoint.
// x: fp32 [B, N, 128] (the LayerNorm output).
#include "common.cuh"

namespace dml {
namespace {

constexpr int kCc = 128;       // channels

// ---------------------------------------------------------------------------------------------------------------------
// Tiled N x N squared distances on the tensor cores: d2(i, j) = |x_i|^2 + |x_j|^2 - 2 x_i . x_j with the dot products as
// mma.sync.m16n8k16 on fp16 PAIRS (hi + lo = 22 bits, 3 MMAs per product: the arithmetic class of the fp32 matmul torch.cdist
// itself uses for N > 25, ClusterMergeNet.py:88) after an exact power-of-two scaling of x that puts max |x| at 2^12 (the lo part
// of anything above 3e-5 of the maximum is then a normal fp16; an mma.sync costs the scheduler ~8.5 issue cycles, so three bf16
// parts = 6 MMAs would double the sweep).  CTA = 128 rows (warp = 16 rows whose A fragments stay in registers for the whole
// sweep: 64 words) x column tiles of 64 tokens streamed through a double-buffered, XOR-swizzled
// shared-memory ring with cp.async; B fragments by ldmatrix; the epilogue (top-5 insertion / masked minimum) runs on the
// accumulator fragments, so only O(N) results leave the SM.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRowsCta = 128, kColsTile = 64;
constexpr int kParts = 2;                                      // fp16 hi, lo
constexpr uint32_t kPlaneBytes = kColsTile * kCc * 2;          // 16 KB: one fp16 part of a column tile
constexpr uint32_t kStageBytes = kParts * kPlaneBytes + 512;   // + |x_j|^2 [64] and density_j [64]
constexpr uint32_t kDistSmem = 2 * kStageBytes;

// planes: fp16 [2][rows][128] = hi / lo parts of x * s, s = the power of two that puts amax (device scalar: max |x|) at 2^12;
// norms [rows] = |x|^2 in fp32 (unscaled); scale_out[0] = 1 / s^2.  One warp per row.
__global__ void __launch_bounds__(256) dpc_split_kernel(const float* __restrict__ x, long long rows, const float* __restrict__ amax,
                                                        uint32_t* __restrict__ planes, float* __restrict__ norms,
                                                        float* __restrict__ scale_out) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const float mx = amax[0];
  const float sc = mx > 0.f ? exp2f(12.f - ceilf(log2f(mx))) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[0] = 1.f / (sc * sc);
  if (r >= rows) return;
  const float4 v = *reinterpret_cast<const float4*>(x + r * kCc + lane * 4);
  uint32_t h0, l0, h1, l1;
  split_f16(v.x * sc, v.y * sc, h0, l0);
  split_f16(v.z * sc, v.w * sc, h1, l1);
  const size_t plane = (size_t)rows * (kCc / 2);             // in 32-bit words
  uint32_t* o = planes + r * (kCc / 2) + lane * 2;
  *reinterpret_cast<uint2*>(o) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(o + plane) = make_uint2(l0, l1);
  const float n2 = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
  if (lane == 0) norms[r] = n2;
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

// kParent = false: density[b, i] = exp(-mean of the 5 smallest d2 / C) + 1e-6 noise, aux_out = row maximum of d2   (:98-104)
// kParent = true : aux_out = parent[b, i] = min(dist_max[b], min_{j: density_j > density_i} sqrt(d2 / C))          (:111-114)
template <bool kParent>
__global__ void __launch_bounds__(256, 1) dpc_dist_kernel(const uint32_t* __restrict__ planes, const float* __restrict__ norms,
                                                           const float* __restrict__ noise, const float* __restrict__ dist_max,
                                                           const float* __restrict__ inv_scale2, int N, float* __restrict__ density,
                                                           float* __restrict__ aux_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y;
  const size_t rows_all = (size_t)gridDim.y * N, plane_w = rows_all * (kCc / 2);
  const size_t brow = (size_t)b * N;
  const int i0 = blockIdx.x * kRowsCta + warp * 16 + g, i1 = i0 + 8;                  // the thread's two rows
  // A fragments of the warp's 16 rows: [part][k-step][reg]
  uint32_t a[kParts][8][4];
#pragma unroll
  for (int p = 0; p < kParts; ++p)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = (r & 1) ? i1 : i0;
        const int colw = 8 * ks + t + ((r & 2) ? 4 : 0);                                // 32-bit word = 2 channels
        a[p][ks][r] = row < N ? planes[p * plane_w + (brow + row) * (kCc / 2) + colw] : 0u;
      }
  const float ni0 = i0 < N ? norms[brow + i0] : 0.f, ni1 = i1 < N ? norms[brow + i1] : 0.f;
  const float m2s = -2.f * inv_scale2[0];                                               // un-scales the dot products
  float di0 = 0.f, di1 = 0.f;
  if (kParent) {
    di0 = i0 < N ? density[brow + i0] : INFINITY;
    di1 = i1 < N ? density[brow + i1] : INFINITY;
  }
  float best0[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY}, best1[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
  float ex0 = kParent ? INFINITY : 0.f, ex1 = ex0;                                      // running min (parent) / max (density)

  const int ntiles = cdiv(N, kColsTile);
  auto load_tile = [&](int tile, int buf) {
    const int j0 = tile * kColsTile;
    unsigned char* base = smem + buf * kStageBytes;
    for (int c = threadIdx.x; c < kParts * kColsTile * 16; c += 256) {
      const int part = c >> 10, row = (c >> 4) & 63, ch = c & 15;
      const bool ok = j0 + row < N;
      const uint32_t* src = planes + part * plane_w + (brow + (ok ? j0 + row : 0)) * (kCc / 2) + ch * 4;
      cp_async16(smem_u32(base + part * kPlaneBytes + row * 256 + ((ch ^ (row & 7)) << 4)), src, ok);
    }
    if (threadIdx.x < kColsTile) {
      const int j = j0 + threadIdx.x;
      float* f = reinterpret_cast<float*>(base + kParts * kPlaneBytes);
      f[threadIdx.x] = j < N ? norms[brow + j] : 0.f;
      if (kParent) f[64 + threadIdx.x] = j < N ? density[brow + j] : -INFINITY;
    }
  };
  load_tile(0, 0);
  cp_async_commit();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const unsigned char* base = smem + buf * kStageBytes;
    const uint32_t sbase = smem_u32(base);
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t bf[kParts][8][2];                             // [part][n-tile][b0, b1]
#pragma unroll
      for (int p = 0; p < kParts; ++p)
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
          const int mi = lane >> 3, row = 16 * ntp + (mi >> 1) * 8 + (lane & 7), ch = 2 * ks + (mi & 1);
          uint32_t r4[4];
          ldmatrix_x4(r4, sbase + p * kPlaneBytes + row * 256 + ((ch ^ (row & 7)) << 4));
          bf[p][2 * ntp][0] = r4[0];
          bf[p][2 * ntp][1] = r4[1];
          bf[p][2 * ntp + 1][0] = r4[2];
          bf[p][2 * ntp + 1][1] = r4[3];
        }
      // (a part, b part): lo.hi, hi.lo, hi.hi; 8 independent accumulators between dependent MMAs
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_f16_16816(acc[nt], a[1][ks], bf[0][nt][0], bf[0][nt][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_f16_16816(acc[nt], a[0][ks], bf[1][nt][0], bf[1][nt][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_f16_16816(acc[nt], a[0][ks], bf[0][nt][0], bf[0][nt][1]);
    }
    // epilogue on the accumulator fragments
    const float* nj = reinterpret_cast<const float*>(base + kParts * kPlaneBytes);
    const int j0 = tile * kColsTile;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = 8 * nt + 2 * t, j = j0 + c;
      const float2 n2 = *reinterpret_cast<const float2*>(nj + c);
      float d00 = fmaxf(fmaf(m2s, acc[nt][0], ni0 + n2.x), 0.f), d01 = fmaxf(fmaf(m2s, acc[nt][1], ni0 + n2.y), 0.f);
      float d10 = fmaxf(fmaf(m2s, acc[nt][2], ni1 + n2.x), 0.f), d11 = fmaxf(fmaf(m2s, acc[nt][3], ni1 + n2.y), 0.f);
      if (j == i0) d00 = 0.f;
      if (j + 1 == i0) d01 = 0.f;
      if (j == i1) d10 = 0.f;
      if (j + 1 == i1) d11 = 0.f;
      if (kParent) {
        const float2 dj = *reinterpret_cast<const float2*>(nj + 64 + c);       // -inf beyond N: never "higher"
        if (dj.x > di0) ex0 = fminf(ex0, d00);
        if (dj.y > di0) ex0 = fminf(ex0, d01);
        if (dj.x > di1) ex1 = fminf(ex1, d10);
        if (dj.y > di1) ex1 = fminf(ex1, d11);
      } else {
        if (j < N) {
          insert5(best0, d00);
          insert5(best1, d10);
          ex0 = fmaxf(ex0, d00);
          ex1 = fmaxf(ex1, d10);
        }
        if (j + 1 < N) {
          insert5(best0, d01);
          insert5(best1, d11);
          ex0 = fmaxf(ex0, d01);
          ex1 = fmaxf(ex1, d11);
        }
      }
    }
    __syncthreads();
  }
  // combine the four threads of a row (columns were split over the quad)
  const float invc = 1.f / sqrtf((float)kCc);
  if (kParent) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      ex0 = fminf(ex0, __shfl_xor_sync(0xffffffffu, ex0, o));
      ex1 = fminf(ex1, __shfl_xor_sync(0xffffffffu, ex1, o));
    }
    if (t == 0) {
      const float dm = dist_max[b];
      if (i0 < N) aux_out[brow + i0] = ex0 == INFINITY ? dm : fminf(dm, sqrtf(ex0) * invc);
      if (i1 < N) aux_out[brow + i1] = ex1 == INFINITY ? dm : fminf(dm, sqrtf(ex1) * invc);
    }
  } else {
    float m0[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY}, m1[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
    for (int src = 0; src < 4; ++src)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        insert5(m0, __shfl_sync(0xffffffffu, best0[k], (lane & ~3) | src));
        insert5(m1, __shfl_sync(0xffffffffu, best1[k], (lane & ~3) | src));
      }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      ex0 = fmaxf(ex0, __shfl_xor_sync(0xffffffffu, ex0, o));
      ex1 = fmaxf(ex1, __shfl_xor_sync(0xffffffffu, ex1, o));
    }
    if (t == 0) {
      // reference: dist = cdist / sqrt(C); density = exp(-mean(dist_nearest^2)) + rand * 1e-6   (ClusterMergeNet.py:88-104)
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float e0 = sqrtf(m0[k]) * invc, e1 = sqrtf(m1[k]) * invc;
        s0 += e0 * e0;
        s1 += e1 * e1;
      }
      if (i0 < N) {
        density[brow + i0] = expf(-(s0 / 5.f)) + noise[brow + i0] * 1e-6f;
        aux_out[brow + i0] = ex0;
      }
      if (i1 < N) {
        density[brow + i1] = expf(-(s1 / 5.f)) + noise[brow + i1] * 1e-6f;
        aux_out[brow + i1] = ex1;
      }
    }
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

// merged[b, c, :] = sum_{i in c} x_i w_i / W_c,  W_c = sum_{i in c} w_i + 1e-6   (:151-166), deterministic and token-parallel:
// a CTA owns a chunk of kMergeChunk tokens and keeps one partial row per cluster in shared memory (thread = channel, so the
// read-modify-writes never collide); the partials [chunk][K][128 + 1] are summed over the chunks in a fixed order afterwards.
constexpr int kMergeChunk = 256;
__global__ void __launch_bounds__(128) merge_part_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const long long* __restrict__ idx, int N, int K, float* __restrict__ parts) {
  extern __shared__ float msm[];
  float* acc = msm;                       // [K][128]
  float* wsum = acc + (size_t)K * kCc;    // [K]
  int* cid = reinterpret_cast<int*>(wsum + K);          // [kMergeChunk]
  float* cw = reinterpret_cast<float*>(cid + kMergeChunk);
  const int b = blockIdx.y, i0 = blockIdx.x * kMergeChunk, k = threadIdx.x;
  for (int i = k; i < K * kCc; i += 128) acc[i] = 0.f;
  for (int i = k; i < K; i += 128) wsum[i] = 0.f;
  const int cnt = min(kMergeChunk, N - i0);
  for (int i = k; i < cnt; i += 128) {
    cid[i] = (int)idx[(size_t)b * N + i0 + i];
    cw[i] = w[(size_t)b * N + i0 + i];
  }
  __syncthreads();
  for (int i = 0; i < cnt; ++i) {
    const int c = cid[i];
    const float wi = cw[i];
    acc[c * kCc + k] += x[((size_t)b * N + i0 + i) * kCc + k] * wi;
    if (k == 0) wsum[c] += wi;
  }
  __syncthreads();
  float* o = parts + ((size_t)b * gridDim.x + blockIdx.x) * K * (kCc + 1);
  for (int i = k; i < K * kCc; i += 128) o[(i / kCc) * (kCc + 1) + (i % kCc)] = acc[i];
  for (int i = k; i < K; i += 128) o[i * (kCc + 1) + kCc] = wsum[i];
}

__global__ void __launch_bounds__(128) merge_final_kernel(const float* __restrict__ parts, int chunks, int K, float* __restrict__ merged,
                                                          float* __restrict__ all_w) {
  const int c = blockIdx.x, b = blockIdx.y, k = threadIdx.x;
  float s = 0.f, W = 0.f;
  for (int ch = 0; ch < chunks; ++ch) {
    const float* p = parts + (((size_t)b * chunks + ch) * K + c) * (kCc + 1);
    s += p[k];
    W += p[kCc];
  }
  W += 1e-6f;
  merged[((size_t)b * K + c) * kCc + k] = s / W;
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

int dml_dpc_split(const float* x, const float* amax, long long rows, int C, void* planes, float* norms, float* inv_scale2, void* stream) {
  DML_CHECK_ARG(x && amax && planes && norms && inv_scale2 && rows > 0);
  if (C != kCc) return DML_EUNSUPPORTED;
  dpc_split_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, amax, reinterpret_cast<uint32_t*>(planes), norms,
                                                                                 inv_scale2);
  DML_RETURN_LAUNCH();
}

int dml_dpc_density(const void* planes, const float* norms, const float* inv_scale2, const float* noise, int B, int N, int C, float* density,
                    float* rowmax2, void* stream) {
  DML_CHECK_ARG(planes && norms && inv_scale2 && noise && density && rowmax2 && B > 0 && N >= 5 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(dpc_dist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDistSmem);
  if (e != cudaSuccess) return (int)e;
  dpc_dist_kernel<false><<<dim3(cdiv(N, kRowsCta), B), 256, kDistSmem, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(planes), norms, noise, nullptr, inv_scale2, N, density, rowmax2);
  DML_RETURN_LAUNCH();
}

int dml_dpc_parent(const void* planes, const float* norms, const float* inv_scale2, const float* density, const float* dist_max, int B, int N,
                   int C, float* parent, void* stream) {
  DML_CHECK_ARG(planes && norms && inv_scale2 && density && dist_max && parent && B > 0 && N > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(dpc_dist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDistSmem);
  if (e != cudaSuccess) return (int)e;
  dpc_dist_kernel<true><<<dim3(cdiv(N, kRowsCta), B), 256, kDistSmem, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(planes), norms, nullptr, dist_max, inv_scale2, N, const_cast<float*>(density), parent);
  DML_RETURN_LAUNCH();
}

int dml_dpc_assign(const float* x, const long long* centres, int B, int N, int C, int K, long long* idx, void* stream) {
  DML_CHECK_ARG(x && centres && idx && B > 0 && N > 0 && K > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  dpc_assign_kernel<<<dim3(cdiv(N, 8), B), 256, 0, (cudaStream_t)stream>>>(x, centres, N, K, idx);
  DML_RETURN_LAUNCH();
}

long long dml_merge_ws_floats(int B, int N, int K) { return (long long)B * cdiv(N, kMergeChunk) * K * (kCc + 1); }

int dml_merge_fwd(const float* x, const float* w, const long long* idx, int B, int N, int C, int K, float* ws, float* merged, float* all_w,
                  void* stream) {
  DML_CHECK_ARG(x && w && idx && ws && merged && all_w && B > 0 && N > 0 && K > 0 && B <= 65535);
  if (C != kCc) return DML_EUNSUPPORTED;
  const size_t smem = ((size_t)K * (kCc + 1) + 2 * kMergeChunk) * sizeof(float);
  if (smem > 200 * 1024) return DML_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int chunks = cdiv(N, kMergeChunk);
  merge_part_kernel<<<dim3(chunks, B), 128, smem, st>>>(x, w, idx, N, K, ws);
  merge_final_kernel<<<dim3(K, B), 128, 0, st>>>(ws, chunks, K, merged, all_w);
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
