// Initial iterate of moore_penrose_iter_pinv (models/NystromAttention.py:20-27) and its adjoint:
//     z0 = x^T / (max_i sum_j |x_ij|  *  max_j sum_i |x_ij|),   both maxima GLOBAL over every matrix of the batch and every head
// (quirk T3 / Q9).  x float [NB, m, m].  The reference runs ~8 small torch kernels forward and ~40 backward per layer for this;
// here: one kernel for the absolute row / column sums, one that writes z0 as a bf16 pair (every CTA re-derives the two global
// maxima from the NB m sums - a few KB), and for the backward one partial-sum kernel plus one that assembles
//     dx_ij = G_ji / D + [i is the arg-max row] dr sign(x_ij) + [j is the arg-max column] dc sign(x_ij) (+ an optional addend),
//     D = r c,  S = sum G_ji x_ij,  dD = -S / D^2,  dr = dD c,  dc = dD r.
#include "common.cuh"

namespace dml {
namespace pv {

constexpr int kThreads = 256;

// sums[0][nb][i] = sum_j |x_ij| (row sums), sums[1][nb][j] = sum_i |x_ij| (column sums).  One CTA per matrix.
__global__ void __launch_bounds__(kThreads)
abs_sums_kernel(const float* __restrict__ x, int NB, int m, float* __restrict__ sums) {
  const int nb = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xm = x + (size_t)nb * m * m;
  for (int j = tid; j < m; j += kThreads) {            // column sums: thread = column, rows in sequence (coalesced); 16 independent
    float s[4] = {0.f, 0.f, 0.f, 0.f};                 // loads in flight per thread, four partial sums
    int i = 0;
    for (; i + 16 <= m; i += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldg(xm + (size_t)(i + u) * m + j);
#pragma unroll
      for (int u = 0; u < 16; ++u) s[u & 3] += fabsf(v[u]);
    }
    for (; i < m; ++i) s[0] += fabsf(__ldg(xm + (size_t)i * m + j));
    sums[((size_t)NB + nb) * m + j] = (s[0] + s[1]) + (s[2] + s[3]);
  }
  for (int i0 = warp * 4; i0 < m; i0 += kThreads / 32 * 4) {      // row sums: a warp takes 4 rows at a time
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < m; j += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u < m) s[u] += fabsf(__ldg(xm + (size_t)(i0 + u) * m + j));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(s[u]);
      if (lane == 0 && i0 + u < m) sums[(size_t)nb * m + i0 + u] = t;
    }
  }
}

struct Maxima { float r, c; int ir, ic; };      // values and flat indices (nb * m + i) of the two global maxima

__device__ __forceinline__ Maxima block_maxima(const float* __restrict__ sums, int total, float* sv, int* si) {
  // every thread scans a strided share of both arrays, then a shared-memory tournament; first index wins ties
  Maxima mx{-1.f, -1.f, 0, 0};
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    const float a = __ldg(sums + k), b = __ldg(sums + total + k);
    if (a > mx.r) { mx.r = a; mx.ir = k; }
    if (b > mx.c) { mx.c = b; mx.ic = k; }
  }
  for (int which = 0; which < 2; ++which) {
    float v = which ? mx.c : mx.r;
    int idx = which ? mx.ic : mx.ir;
    sv[threadIdx.x] = v; si[threadIdx.x] = idx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const float v2 = sv[threadIdx.x + o];
        const int i2 = si[threadIdx.x + o];
        if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) { sv[threadIdx.x] = v2; si[threadIdx.x] = i2; }
      }
      __syncthreads();
    }
    if (which) { mx.c = sv[0]; mx.ic = si[0]; } else { mx.r = sv[0]; mx.ir = si[0]; }
    __syncthreads();
  }
  return mx;
}

// z0[nb][j][i] = x[nb][i][j] / (r c) as a bf16 pair; 32 x 32 tiles transposed through shared memory
__global__ void __launch_bounds__(kThreads)
z0_kernel(const float* __restrict__ x, const float* __restrict__ sums, int NB, int m, bf16* __restrict__ z, long long plane) {
  __shared__ float sv[kThreads];
  __shared__ int si[kThreads];
  __shared__ float tile[32][33];
  const Maxima mx = block_maxima(sums, NB * m, sv, si);
  const float inv = 1.0f / (mx.r * mx.c);
  const int tiles = (m + 31) / 32;
  const int nb = blockIdx.y, t = blockIdx.x, ti = t / tiles, tj = t % tiles;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const float* xm = x + (size_t)nb * m * m;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = ti * 32 + r, j = tj * 32 + tx;
    tile[r][tx] = (i < m && j < m) ? __ldg(xm + (size_t)i * m + j) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int j = tj * 32 + r, i = ti * 32 + tx;               // output row j, column i
    if (i < m && j < m) {
      const float v = tile[tx][r] / (mx.r * mx.c);             // the reference divides (x^T / denom): same rounding
      const bf16 hi = __float2bfloat16(v);
      const size_t o = ((size_t)nb * m + j) * m + i;
      z[o] = hi;
      z[plane + o] = __float2bfloat16(v - __bfloat162float(hi));
    }
  }
  (void)inv;
}

// part[cta] = sum over the CTA's elements of G[nb][j][i] x[nb][i][j]
__global__ void __launch_bounds__(kThreads)
gx_dot_kernel(const float* __restrict__ G, const float* __restrict__ x, int m, float* __restrict__ part) {
  __shared__ float tile[32][33];
  __shared__ float red[kThreads / 32];
  const int tiles = (m + 31) / 32;
  const int nb = blockIdx.y, t = blockIdx.x, ti = t / tiles, tj = t % tiles;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xm = x + (size_t)nb * m * m;
  const float* gm = G + (size_t)nb * m * m;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {                           // G tile (rows tj.., cols ti..) -> shared, read back transposed
    const int j = tj * 32 + r, i = ti * 32 + tx;
    tile[r][tx] = (i < m && j < m) ? __ldg(gm + (size_t)j * m + i) : 0.f;
  }
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = ti * 32 + r, j = tj * 32 + tx;
    if (i < m && j < m) s = fmaf(tile[tx][r], __ldg(xm + (size_t)i * m + j), s);
  }
  s = warp_sum(s);
  if (tx == 0) red[ty] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
    part[(size_t)nb * gridDim.x + t] = tot;
  }
}

__global__ void __launch_bounds__(kThreads)
dx_kernel(const float* __restrict__ G, const float* __restrict__ x, const float* __restrict__ sums, const float* __restrict__ part,
          int npart, const float* __restrict__ addend, int NB, int m, float* __restrict__ dx) {
  __shared__ float sv[kThreads];
  __shared__ int si[kThreads];
  __shared__ float tile[32][33];
  __shared__ float s_S;
  const Maxima mx = block_maxima(sums, NB * m, sv, si);
  float ps = 0.f;                                              // the same summation order in every CTA: bit-identical S
  for (int k = threadIdx.x; k < npart; k += kThreads) ps += __ldg(part + k);
  sv[threadIdx.x] = ps;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sv[threadIdx.x] += sv[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_S = sv[0];
  __syncthreads();
  const float D = mx.r * mx.c;
  const float dD = -s_S / (D * D), dr = dD * mx.c, dc = dD * mx.r;
  const int tiles = (m + 31) / 32;
  const int nb = blockIdx.y, t = blockIdx.x, ti = t / tiles, tj = t % tiles;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* gm = G + (size_t)nb * m * m;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int j = tj * 32 + r, i = ti * 32 + tx;
    tile[r][tx] = (i < m && j < m) ? __ldg(gm + (size_t)j * m + i) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = ti * 32 + r, j = tj * 32 + tx;
    if (i < m && j < m) {
      const size_t o = ((size_t)nb * m + i) * m + j;
      const float xv = __ldg(x + o);
      const float sg = xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f);
      float v = tile[tx][r] / D;
      if (nb * m + i == mx.ir) v = fmaf(dr, sg, v);
      if (nb * m + j == mx.ic) v = fmaf(dc, sg, v);
      if (addend != nullptr) v += __ldg(addend + o);
      dx[o] = v;
    }
  }
}

}  // namespace pv
}  // namespace dml

extern "C" {

/* floats of the sums workspace (saved for the backward) and of the backward's partial-sum workspace */
size_t dml_ny_pinv_init_sums_floats(int NB, int m) { return NB > 0 && m > 0 ? (size_t)2 * NB * m : 0; }
size_t dml_ny_pinv_init_part_floats(int NB, int m) { return NB > 0 && m > 0 ? (size_t)NB * ((m + 31) / 32) * ((m + 31) / 32) : 0; }

int dml_ny_pinv_init_fwd(const float* x, int NB, int m, float* sums, void* z_pair, long long plane_stride, void* stream) {
  using namespace dml;
  DML_CHECK_ARG(x && sums && z_pair && NB > 0 && m > 0 && NB <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  pv::abs_sums_kernel<<<NB, pv::kThreads, 0, st>>>(x, NB, m, sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int tiles = (m + 31) / 32;
  pv::z0_kernel<<<dim3(tiles * tiles, NB), pv::kThreads, 0, st>>>(x, sums, NB, m, (bf16*)z_pair, plane_stride);
  DML_RETURN_LAUNCH();
}

int dml_ny_pinv_init_bwd(const float* g, const float* x, const float* sums, const float* addend, int NB, int m, float* part, float* dx,
                         void* stream) {
  using namespace dml;
  DML_CHECK_ARG(g && x && sums && part && dx && NB > 0 && m > 0 && NB <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (m + 31) / 32;
  dim3 grid(tiles * tiles, NB);
  pv::gx_dot_kernel<<<grid, pv::kThreads, 0, st>>>(g, x, m, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  pv::dx_kernel<<<grid, pv::kThreads, 0, st>>>(g, x, sums, part, tiles * tiles * NB, addend, NB, m, dx);
  DML_RETURN_LAUNCH();
}

}  // extern "C"
